"""One streaming-shaped step (BASELINE config 5: B x 9 600 samples, separation + scoring of both streams) launched
eagerly (CUDA graphs off) inside a cudaProfilerStart/Stop range - the ncu target for the per-kernel launch list of the
batch-1 step:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
      python tools/run_stream_once.py [B = 1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
stage = SeparationScoringStage.random_init("cuda:0", seed=0)
stage.separator.graph_max_frames = 0
target = stage.embed(synthetic_mixture(1, 32000, seed=99).cuda())[0]
mix = synthetic_mixture(B, 9600, seed=5).cuda()
for _ in range(3):
    stage.run(mix, target)
torch.cuda.synchronize()
torch.cuda.profiler.start()
stage.run(mix, target)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one streaming step, batch", B)
