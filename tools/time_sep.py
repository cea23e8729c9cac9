"""Times the separator alone on the GPU (development aid): python tools/time_sep.py [B] [T] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200.synth import random_state_dict, synthetic_mixture  # noqa: E402
from targetdiarization_b200 import Separator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sep = Separator(random_state_dict(0), "cuda:0")
mix = synthetic_mixture(B, T).cuda()
for _ in range(2):
    sep(mix)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sep(mix)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps(dict(B=B, T=T, ms=ms, xrt=B * T / 16000 / (ms / 1e3), ws_gb=sep.workspace_bytes(B, T) / 1e9)))
if os.environ.get("TDZ_STEPS", "1") == "1":
    tab = sep.time_steps(mix, reps=3)
    layer = list(sep.LAYER_STEPS)
    tot = sum(v * (24 if n in layer else 1) for n, v in tab.items())
    for n, v in sorted(tab.items(), key=lambda kv: -kv[1] * (24 if kv[0] in layer else 1)):
        m = 24 if n in layer else 1
        print(f"{n:10s} {v:8.3f} ms x{m:2d} = {v * m:8.2f} ms  {100 * v * m / tot:5.1f}%")
    print("sum of steps", tot)
