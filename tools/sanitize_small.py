"""Small end-to-end run for compute-sanitizer (memcheck): one ragged separation, overlap-add, scoring, loudness.
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import SeparationScoringStage  # noqa: E402
from targetdiarization_b200.synth import synthetic_mixture  # noqa: E402

stage = SeparationScoringStage.random_init("cuda:0", seed=0)
mix = synthetic_mixture(2, 4013, seed=1).cuda()      # ragged: S = 500, tail of 5 samples
tgt = stage.embed(synthetic_mixture(1, 8000, seed=2).cuda())[0]
est, scores = stage.run(mix, tgt)
s1, s2 = stage.separate_speaker(synthetic_mixture(1, 9000, seed=3)[0].numpy())
y = stage.wav_chunk_inference(synthetic_mixture(1, 2500, seed=4)[None].cuda(), sr=100)   # 12 s = 1200 samples sessions
torch.cuda.synchronize()
print("ok", tuple(est.shape), scores.tolist(), s1.shape, tuple(y.shape))
