"""Times the speaker-scoring path alone on the GPU (development aid): python tools/time_sv.py [N] [T] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200.embedder import Embedder  # noqa: E402
from targetdiarization_b200.synth import random_eres2netv2_state_dict, synthetic_mixture  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
emb = Embedder(random_eres2netv2_state_dict(0), "cuda:0")
wav = synthetic_mixture(N, T, seed=3).cuda()
for _ in range(2):
    emb.embed_many(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    emb.embed_many(wav)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps(dict(N=N, T=T, ms=ms, stream_seconds_per_second=N * T / 16000 / (ms / 1e3), sub_batch=emb.max_batch(398))))
