import os, sys
sys.path.insert(0, '/root/repo')
import torch
from targetdiarization_b200 import Separator
from targetdiarization_b200.synth import random_state_dict
B, T = 2, 9613
sd = random_state_dict(seed=1)
g = torch.Generator().manual_seed(1235)
mix = (torch.randn(B, T, generator=g) * 0.1).cuda()
sep = Separator(sd, "cuda:0")
def snr(a, b):
    a = a.double(); b = b.double()
    return float(10 * torch.log10((a ** 2).sum() / ((a - b) ** 2).sum().clamp(min=1e-300)))
lay = sep.layout(B, T)
S = lay.S
for nl in (1, 2, 3, 24):
    res = {}
    for mode in ("old", "new"):
        if mode == "old": os.environ["TDZ_CONV_ROWMAJOR"] = "1"
        else: os.environ.pop("TDZ_CONV_ROWMAJOR", None)
        out = sep(mix, _debug=(nl, 0, 20)).clone()
        torch.cuda.synchronize()
        x = sep.debug_buffer(B, T, "x", torch.float32, 512)[:, :S].clone()
        res[mode] = (out, x)
    o0, x0 = res["old"]; o1, x1 = res["new"]
    print(nl, "out snr", snr(o0, o1), "x snr b0", snr(x0[0], x1[0]), "b1", snr(x0[1], x1[1]), "nan", bool(torch.isnan(x1).any()), float(o1.abs().max()))
