"""Stand-in for the un-pinned third-party `rotary_embedding_torch` package (lucidrains), which the reference
imports at look2hear/models/mossformer_block.py:6 and which is not installed in this image.

TEST INFRASTRUCTURE ONLY (oracle).  Restates the published algorithm used by the reference's call sites
(`RotaryEmbedding(dim=32)` at mossformer_block.py:453 and `rotate_queries_or_keys` at :230-233):
  freqs_j = 1 / 10000^(2j/dim), j < dim/2 (a parameter, so it appears in the state dict);
  angle[t, 2j] = angle[t, 2j+1] = t * freqs_j with t = 0..n-1 along dim -2;
  first `dim` features: x*cos + rotate_half(x)*sin with rotate_half on interleaved pairs (x0,x1)->(-x1,x0);
  remaining features pass through.
Parity unpinned: the reference holds no golden vector for this dependency.
"""
import torch
from torch import nn


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)
        self.dim = dim

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device, dtype=self.freqs.dtype)
        ang = torch.einsum("i,j->ij", pos, self.freqs)
        ang = ang.repeat_interleave(2, dim=-1)  # [n, dim]
        rot_dim = ang.shape[-1]
        t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
        t_rot = t_rot * ang.cos() + _rotate_half(t_rot) * ang.sin()
        return torch.cat((t_rot, t_pass), dim=-1)
