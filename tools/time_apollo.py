"""Times the Apollo restorer (tdz_apollo_restore) and the MDX STFT / iSTFT pair on cuda:0 with CUDA events.
usage: python tools/time_apollo.py [seconds of 44.1 kHz audio = 10] [rows = 1]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from targetdiarization_b200 import ConvTDFNet, Restorer, synth  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ns = int(secs * 44100)
    rest = Restorer(synth.random_apollo_state_dict(0), "cuda:0")
    x = synth.synthetic_fullband(rows, ns, seed=1).reshape(rows, 1, ns).cuda()
    ms = timed(lambda: rest(x))
    T = 1 + ns // 441
    tokens = rows * T * 80
    flop = tokens * 6 * 2 * (256 * 768 + 256 * 256 + 256 * 2048 + 1024 * 256 + 3 * 2 * 256 * 1024)
    rec = {"apollo_ms": ms, "audio_s": secs * rows, "x_realtime": secs * rows / (ms / 1e3), "tokens": tokens,
           "gemm_tflops": flop / (ms * 1e-3) / 1e12}
    net = ConvTDFNet("vocals", 11, 3072, 8, 6144, 1024, "cuda:0")
    xm = torch.randn(8, 2, net.chunk_size, device="cuda") * 0.1
    spec = net.stft(xm)
    rec["mdx_stft_ms_per_8_chunks"] = timed(lambda: net.stft(xm))
    rec["mdx_istft_ms_per_8_chunks"] = timed(lambda: net.istft(spec))   # includes the reference's .cpu()
    rec["mdx_audio_s"] = 8 * net.chunk_size / 44100
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
