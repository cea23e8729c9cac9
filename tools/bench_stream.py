"""Streaming-shaped benchmark (BASELINE.json config 5): 600 ms chunks (9 600 samples -> 1 199 frames, 5 attention
groups) through separation + scoring of both streams, at batch 1 (latency of one stream) and batch 256 (throughput
of 256 concurrent streams).  The reference processes each incoming chunk with batch 1 (TargetDiarizationStream.py:
189-258); batching concurrent streams into one call is what the device path adds.

  python tools/bench_stream.py [--steps 100]    -> one JSON line"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()
T = 9600
stage = SeparationScoringStage.random_init("cuda:0", seed=0)
target = stage.embed(synthetic_mixture(1, 32000, seed=99).cuda())[0]
out = {"chunk_samples": T, "chunk_ms": 600.0, "steps": a.steps}
for B in (1, 256):
    mix_host = synthetic_mixture(B, T, seed=5).pin_memory()
    scores_host = torch.empty(B, 2).pin_memory()
    est_host = torch.empty(B, 2, T).pin_memory()
    lat = []
    for i in range(a.steps + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m = mix_host.to("cuda:0", non_blocking=True)
        est, scores = stage.run(m, target)
        est_host.copy_(est, non_blocking=True)
        scores_host.copy_(scores, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            lat.append(e0.elapsed_time(e1))
    lat.sort()
    p50, p99 = lat[len(lat) // 2], lat[min(len(lat) - 1, int(len(lat) * 0.99))]
    out[f"batch{B}"] = dict(p50_ms=p50, p99_ms=p99, stream_seconds_per_second=B * 0.6 / (p50 / 1e3),
                            realtime_factor_per_stream=0.6 / (p50 / 1e3))
print(json.dumps(out))
