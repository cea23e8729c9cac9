"""SNR of tiny inputs over several seeds (development aid): python tools/tiny_snr.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from oracle.mossformer2_port import mossformer2_forward, snr_db  # noqa: E402
from targetdiarization_b200.synth import random_state_dict  # noqa: E402
from targetdiarization_b200 import Separator  # noqa: E402

for seed in (4, 5, 6):
    sd = random_state_dict(seed=seed)
    sep = Separator(sd, "cuda:0")
    for T in (16, 23, 40, 100):
        vals = []
        for ds in range(4):
            g = torch.Generator().manual_seed(1234 + seed + 100 * ds)
            mix = torch.randn(1, T, generator=g) * 0.1
            with torch.no_grad():
                ref = mossformer2_forward(sd, mix)
            vals.append(snr_db(ref, sep(mix.cuda()).cpu()))
        print(f"weights seed {seed} T={T}: " + " ".join(f"{v:.1f}" for v in vals))
    del sep
