"""CPU restatement of the reference's Apollo restorer (look2hear/models/apollo.py) and of the MDX-Net STFT / iSTFT
front / back end (AudioProcessor.py:65-120): SURVEY.md section 8f-4.

TEST INFRASTRUCTURE ONLY (oracle): imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by
the product.  Pinned by the reference module itself (tests/test_apollo_oracle.py, where /root/reference exists) and
by golden vectors produced by running that module (tests/golden/apollo_small.npz, oracle/make_golden.py).

The restatement is functional (a state dict in, no nn.Module) and TOKEN-MAJOR like the CUDA path: activations are
[B', T, nband, N] instead of the reference's [B', nband, N, T], so every tap compares one to one with a device buffer.
"""
import math

import torch
import torch.nn.functional as F

EPS_RMS = 1e-5


def band_widths(win):
    """apollo.py:232-236"""
    enc_dim = win // 2 + 1
    bw = [int(win / 160)] * 79
    bw.append(enc_dim - sum(bw))
    return bw


def rmsnorm(x, w):
    """apollo.py:7-23 with groups = 1, channels last."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + EPS_RMS) * w


def _q(x, operands):
    """Emulates the rounding of a tensor-core operand ('bf16' / 'fp16' / None = fp32)."""
    if operands == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if operands == "fp16":
        return x.to(torch.float16).to(torch.float32)
    return x


def lin(x, w, b=None, operands=None):
    y = _q(x, operands) @ _q(w.squeeze(-1), operands).t()
    return y if b is None else y + b


def stft(x, n_fft, hop):
    """torch.stft as the reference calls it (center=True, reflect padding, periodic Hann): [R, L] -> [R, T, bins] complex
    (frame-major).  apollo.py:261-262, AudioProcessor.py:85-92."""
    w = torch.hann_window(n_fft, periodic=True, dtype=x.dtype)
    return torch.stft(x, n_fft=n_fft, hop_length=hop, window=w, center=True, return_complex=True).transpose(1, 2)


def istft(spec, n_fft, hop, length=None):
    """[R, T, bins] complex -> [R, L]  (apollo.py:294-295, AudioProcessor.py:116-118)."""
    w = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    return torch.istft(spec.transpose(1, 2), n_fft=n_fft, hop_length=hop, window=w, center=True, length=length)


def band_split(sd, spec, bw):
    """apollo.py:256-293 (spec_band_split + feature_extractor): [R, T, 442] complex -> [R, T, nband, N]."""
    eps = torch.finfo(torch.float32).eps
    feats = []
    k0 = 0
    for i, w in enumerate(bw):
        s = spec[..., k0:k0 + w]
        power = (s.abs().pow(2).sum(-1, keepdim=True) + eps).sqrt()
        f = torch.cat([s.real / power, s.imag / power, torch.log(power)], -1)
        f = rmsnorm(f, sd[f"BN.{i}.0.weight"])
        feats.append(f @ sd[f"BN.{i}.1.weight"].squeeze(-1).t() + sd[f"BN.{i}.1.bias"])
        k0 += w
    return torch.stack(feats, 2)


def rotary(x, cos, sin):
    """apollo.py:107-118: pairs (a, b) -> (a cos - b sin, b cos + a sin); x [..., seq, hd]."""
    seq = x.shape[-2]
    a, b = x[..., 0::2], x[..., 1::2]
    neg = torch.stack((-b, a), -1).reshape(x.shape)
    return x * cos[:seq] + neg * sin[:seq]


def roformer(sd, p, x, operands=None, taps=None):
    """apollo.py:120-141 on x [n_seq, seq, N] (non-causal, 8 heads)."""
    n_seq, seq, N = x.shape
    H, hd = 8, N // 8
    qkv = lin(rmsnorm(x, 1.0), sd[p + "weight.weight"] * sd[p + "input_norm.weight"][None, :, None], operands=operands)
    qkv = qkv.reshape(n_seq, seq, H, 3 * hd).transpose(1, 2)            # [n_seq, H, seq, 3 hd]
    q, k, v = qkv.split(hd, dim=-1)
    q = rotary(q, sd[p + "cos_freq"], sd[p + "sin_freq"])
    k = rotary(k, sd[p + "cos_freq"], sd[p + "sin_freq"])
    att = F.scaled_dot_product_attention(q, k, v)                       # [n_seq, H, seq, hd]
    att = att.transpose(1, 2).reshape(n_seq, seq, N)
    if taps is not None:
        taps["att"] = att
    out = lin(att, sd[p + "output.weight"], operands=operands) + x
    h = F.silu(lin(rmsnorm(out, 1.0), sd[p + "MLP.1.weight"] * sd[p + "MLP.0.weight"][None, :, None], operands=operands))
    gate, z = h.chunk(2, dim=-1)
    return out + lin(F.silu(gate) * z, sd[p + "MLP_output.weight"], operands=operands)


def conv_act_norm(sd, q, x, operands=None):
    """apollo.py:143-170 (non-causal) on x [n, T, N]: x + conv1x1(SiLU(conv1x1(RMSNorm(dwconv7(x)))))."""
    y = F.conv1d(x.transpose(1, 2), sd[q + "0.weight"], sd[q + "0.bias"], padding=3, groups=x.shape[-1]).transpose(1, 2)
    y = lin(rmsnorm(y, 1.0), sd[q + "2.weight"] * sd[q + "1.weight"][None, :, None], sd[q + "2.bias"], operands)
    return x + lin(F.silu(y), sd[q + "4.weight"], sd[q + "4.bias"], operands)


def bsnet(sd, l, x, operands=None, taps=None):
    """apollo.py:186-212 on x [R, T, nband, N]."""
    R, T, nb, N = x.shape
    y = roformer(sd, f"net.{l}.band_net.", x.reshape(R * T, nb, N), operands, taps).reshape(R, T, nb, N)
    if taps is not None:
        taps["band"] = y
    z = y.transpose(1, 2).reshape(R * nb, T, N)
    for b in range(3):
        z = conv_act_norm(sd, f"net.{l}.seq_net.blocks.{b}.conv.", z, operands)
    return z.reshape(R, nb, T, N).transpose(1, 2)


def band_merge(sd, x, bw):
    """apollo.py:287-293: per band RMSNorm -> Conv1d(N, 4 BW) -> GLU -> (real | imag): [R, T, nband, N] -> [R, T, bins]."""
    out = []
    for i, w in enumerate(bw):
        y = rmsnorm(x[:, :, i], sd[f"output.{i}.0.weight"]) @ sd[f"output.{i}.1.weight"].squeeze(-1).t() \
            + sd[f"output.{i}.1.bias"]
        y = y[..., :2 * w] * torch.sigmoid(y[..., 2 * w:])
        out.append(torch.complex(y[..., :w], y[..., w:]))
    return torch.cat(out, -1)


def apollo_forward(sd, wav, sr=44100, win_ms=20, layer=6, operands=None, taps=None):
    """Apollo.forward (apollo.py:278-297): wav [B, nch, nsample] -> [B, nch, nsample].  `taps`: dict filled with
    intermediates in the device layout (spec, feat, per-layer outputs, est_spec)."""
    B, nch, ns = wav.shape
    win = int(sr * win_ms // 1000)
    hop = win // 2
    bw = band_widths(win)
    spec = stft(wav.reshape(B * nch, ns), win, hop)
    x = band_split(sd, spec, bw)
    if taps is not None:
        taps["spec"] = spec
        taps["feat"] = x
    for l in range(layer):
        lt = {} if taps is not None and l == 0 else None
        x = bsnet(sd, l, x, operands, lt)
        if taps is not None:
            taps[f"layer{l}"] = x
            if lt:
                taps["att0"] = lt["att"]
                taps["band0"] = lt["band"]
    est = band_merge(sd, x, bw)
    if taps is not None:
        taps["est_spec"] = est
    return istft(est, win, hop, length=ns).reshape(B, nch, ns)


# ------------------------------------------------------------------------------------------------ MDX-Net STFT / iSTFT
def mdx_stft(x, n_fft, hop, dim_f):
    """ConvTDFNet.stft (AudioProcessor.py:82-99): x [B, 2, chunk] -> [B, 4, dim_f, dim_t]."""
    chunk = x.shape[-1]
    n_bins = n_fft // 2 + 1
    w = torch.hann_window(n_fft, periodic=True)
    s = torch.stft(x.reshape(-1, chunk), n_fft=n_fft, hop_length=hop, window=w, center=True, return_complex=True)
    s = torch.view_as_real(s).permute(0, 3, 1, 2)
    dim_t = s.shape[-1]
    s = s.reshape(-1, 2, 2, n_bins, dim_t).reshape(-1, 4, n_bins, dim_t)
    return s[:, :, :dim_f].contiguous()


def mdx_istft(x, n_fft, hop):
    """ConvTDFNet.istft (AudioProcessor.py:101-120), target_name != '*': x [B, 4, dim_f, dim_t] -> [B, 2, chunk]."""
    n_bins = n_fft // 2 + 1
    B, _, dim_f, dim_t = x.shape
    x = torch.cat([x, torch.zeros(B, 4, n_bins - dim_f, dim_t)], -2)
    x = x.reshape(-1, 2, n_bins, dim_t).permute(0, 2, 3, 1).contiguous()
    w = torch.hann_window(n_fft, periodic=True)
    y = torch.istft(torch.view_as_complex(x), n_fft=n_fft, hop_length=hop, window=w, center=True)
    return y.reshape(-1, 2, hop * (dim_t - 1))


def snr_db(ref, est):
    ref, est = ref.double(), est.double()
    return float(10 * torch.log10(ref.pow(2).sum() / (ref - est).pow(2).sum().clamp(min=1e-300)))
