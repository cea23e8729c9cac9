"""Runs every launch step of the separator once (layer-0 instance, populated buffers), then one fbank + embedder call,
inside a cudaProfilerStart/Stop range - the target of the per-kernel ncu capture:

  ncu --metrics <list, see tools/ncu_steps.sh> --clock-control none --profile-from-start off -o gpurun_out/steps \
      python tools/run_all_steps.py [B] [T] [N_embed]
(keep B moderate: ncu saves / restores the touched device memory around every replay pass - with the 13 GB workspace of
the 64-item batch and `--set full` the capture of ~220 launches did not finish in 25 minutes)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
NE = int(sys.argv[3]) if len(sys.argv) > 3 else 128
stage = SeparationScoringStage.random_init("cuda:0", seed=0)
sep = stage.separator
mix = synthetic_mixture(B, T).cuda()
est = sep(mix)
wav = est.view(2 * B, T)[:NE].contiguous()
stage.embed(wav)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for k, name in enumerate(sep.STEP_NAMES):
    sep(mix, _debug=(1, k, k))
torch.cuda.synchronize()
stage.embed(wav)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("steps:", " ".join(sep.STEP_NAMES), "| embedder on", NE, "utterances")
