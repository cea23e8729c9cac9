"""Runs ONE launch step of the separator `reps` times on populated buffers (for ncu / timing of a single kernel):
   python tools/run_step.py STEP [B] [T] [reps]        STEP = a name of Separator.STEP_NAMES, e.g. FLASH_IN"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import Separator  # noqa: E402
from targetdiarization_b200.synth import random_state_dict, synthetic_mixture  # noqa: E402

step = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sep = Separator(random_state_dict(0), "cuda:0")
mix = synthetic_mixture(B, T).cuda()
sep(mix)
torch.cuda.synchronize()
k = sep.STEP_NAMES.index(step)
sep(mix, _debug=(1, k, k))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sep(mix, _debug=(1, k, k))
e1.record()
torch.cuda.synchronize()
print(f"{step} B={B} T={T}: {e0.elapsed_time(e1) / reps:.4f} ms")
