"""Random-init weights of the two named architectures and synthetic mixtures (used by bench.py, smoke() and the
tests when no checkpoint is available; BASELINE.json: "random-init of the named architecture when checkpoints are
absent").  State-dict key names and shapes are the reference's (MossFormer2: SURVEY.md section 8a13, verified
key-for-key against the reference module by tests/test_oracle_port.py; ERes2NetV2: the published modelscope /
3D-Speaker module names), so a real checkpoint loads through the same packers.  `perturb=True` additionally
randomises every norm gain/bias, PReLU slope and ScaleNorm g, so that the weight folding of the packer is exercised.
"""
import math

import torch

LAYER = "mask_net.mdl.intra_mdl.mossformerM.layers.{}."
FSMN = "mask_net.mdl.intra_mdl.mossformerM.fsmn.{}."


def _uniform(gen, shape, fan_in):
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def random_state_dict(seed=0, perturb=False, num_layers=24):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, extra=()):
        sd[name + ".weight"] = _uniform(g, (out_f, in_f) + tuple(extra), in_f * (math.prod(extra) if extra else 1))
        if bias:
            sd[name + ".bias"] = _uniform(g, (out_f,), in_f * (math.prod(extra) if extra else 1))

    def affine(name, n, wkey="weight", bkey="bias"):
        if perturb:
            sd[f"{name}.{wkey}"] = 1.0 + 0.2 * torch.randn(n, generator=g)
            if bkey:
                sd[f"{name}.{bkey}"] = 0.1 * torch.randn(n, generator=g)
        else:
            sd[f"{name}.{wkey}"] = torch.ones(n)
            if bkey:
                sd[f"{name}.{bkey}"] = torch.zeros(n)

    def prelu(name, n):
        sd[name] = torch.full((n,), 0.25) + (0.05 * torch.randn(n, generator=g) if perturb else 0.0)

    sd["enc.conv1d.weight"] = _uniform(g, (512, 1, 16), 16)
    affine("mask_net.norm", 512)
    sd["mask_net.conv1d_encoder.weight"] = _uniform(g, (512, 512, 1), 512)
    sd["mask_net.pos_enc.scale"] = torch.ones(1) + (0.1 * torch.randn(1, generator=g) if perturb else 0.0)
    sd["mask_net.pos_enc.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, 512, 2).float() / 512))
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2)[:16].float() / 32))
    for i in range(num_layers):
        q = FSMN.format(i)
        sd[q + "conv1.0.weight"] = _uniform(g, (256, 512, 1), 512)
        sd[q + "conv1.0.bias"] = _uniform(g, (256,), 512)
        prelu(q + "conv1.1.weight", 1)
        affine(q + "norm1", 256)
        for n in ("to_u", "to_v"):
            affine(q + f"gated_fsmn.{n}.mdl.0", 256)
            lin(q + f"gated_fsmn.{n}.mdl.1", 256, 256)
            sd[q + f"gated_fsmn.{n}.mdl.3.sequential.1.conv.weight"] = _uniform(g, (256, 1, 17), 17)
        lin(q + "gated_fsmn.fsmn.linear", 256, 256)
        lin(q + "gated_fsmn.fsmn.project", 256, 256, bias=False)
        c = q + "gated_fsmn.fsmn.conv."
        sd[c + "conv1.weight"] = _uniform(g, (256, 1, 39, 1), 39)
        affine(c + "norm1", 256)
        prelu(c + "prelu1.weight", 256)
        sd[c + "conv2.weight"] = _uniform(g, (256, 2, 39, 1), 78)
        affine(c + "norm2", 256)
        prelu(c + "prelu2.weight", 256)
        affine(q + "norm2", 256)
        sd[q + "conv2.weight"] = _uniform(g, (512, 256, 1), 256)
        sd[q + "conv2.bias"] = _uniform(g, (512,), 256)
    for i in range(num_layers):
        p = LAYER.format(i)
        sd[p + "rotary_pos_emb.freqs"] = freqs.clone()
        for n, (o, k) in (("to_hidden", (2048, 512)), ("to_qk", (128, 512)), ("to_out", (512, 1024))):
            affine(p + f"{n}.mdl.0", 1, wkey="g", bkey=None)
            lin(p + f"{n}.mdl.1", o, k)
            sd[p + f"{n}.mdl.3.sequential.1.conv.weight"] = _uniform(g, (o, 1, 17), 17)
        sd[p + "qk_offset_scale.gamma"] = 0.02 * torch.randn(4, 128, generator=g)
        sd[p + "qk_offset_scale.beta"] = (0.02 * torch.randn(4, 128, generator=g) if perturb
                                          else torch.zeros(4, 128))
    affine("mask_net.mdl.intra_mdl.norm", 512)
    affine("mask_net.mdl.intra_norm", 512)
    sd["mask_net.conv1d_out.weight"] = _uniform(g, (1024, 512, 1), 512)
    sd["mask_net.conv1d_out.bias"] = _uniform(g, (1024,), 512)
    sd["mask_net.conv1_decoder.weight"] = _uniform(g, (512, 512, 1), 512)
    prelu("mask_net.prelu.weight", 1)
    for n in ("output", "output_gate"):
        sd[f"mask_net.{n}.0.weight"] = _uniform(g, (512, 512, 1), 512)
        sd[f"mask_net.{n}.0.bias"] = _uniform(g, (512,), 512)
    sd["dec.weight"] = _uniform(g, (512, 1, 16), 16)
    return sd


def synthetic_mixture(n_items, n_samples, seed=1234, sr=16000):
    """Two independent 'talkers': low-passed white noise under a 4 Hz on/off envelope, summed, peak 0.5
    (SURVEY.md section 8d).  Returns float32 [n_items, n_samples]."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float32) / sr
    out = torch.zeros(n_items, n_samples)
    for s in range(2):
        x = torch.randn(n_items, n_samples, generator=g)
        # one-pole low-pass pair via cumulative filtering in the frequency domain (cheap and deterministic)
        X = torch.fft.rfft(x, dim=-1)
        f = torch.fft.rfftfreq(n_samples, 1.0 / sr)
        fc = 800.0 + 1200.0 * s
        X = X / (1.0 + (f / fc) ** 2)
        x = torch.fft.irfft(X, n=n_samples, dim=-1)
        phase = torch.rand(n_items, 1, generator=g) * 2 * math.pi
        env = (torch.sin(2 * math.pi * 4.0 * t[None, :] * (0.5 + 0.25 * s) + phase) > -0.2).float()
        out += x * env
    out = out / out.abs().amax(dim=-1, keepdim=True).clamp(min=1e-9) * 0.5
    return out.contiguous()


# ------------------------------------------------------------------------------------------------ ERes2NetV2-Large
NUM_BLOCKS = (3, 4, 6, 3)
M_CHANNELS = 64
BASE_WIDTH = 24
SCALE = 4
EXPANSION = 4
FEAT_DIM = 80
EMBED_DIM = 192


def block_specs():
    """[(name, in_planes, planes, stride, is_aff)] for the 16 residual blocks."""
    specs = []
    in_planes = M_CHANNELS
    for li, (n, mult, stride, aff) in enumerate(zip(NUM_BLOCKS, (1, 2, 4, 8), (1, 2, 2, 2), (False, False, True, True))):
        planes = M_CHANNELS * mult
        for bi in range(n):
            specs.append((f"layer{li + 1}.{bi}", in_planes, planes, stride if bi == 0 else 1, aff))
            in_planes = planes * EXPANSION
    return specs


def random_eres2netv2_state_dict(seed=0):
    """Random-init weights with non-trivial BatchNorm running statistics (so BN folding is exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, bias=False):
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)
        # kaiming-uniform(a=sqrt(5)) as nn.Conv2d does
        sd[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 1.0 + 0.2 * torch.rand(c, generator=g)

    def aff(name, channels, r=4):
        inter = channels // r
        conv(name + ".local_att.0", inter, channels * 2, 1, bias=True)
        bn(name + ".local_att.1", inter)
        conv(name + ".local_att.3", channels, inter, 1, bias=True)
        bn(name + ".local_att.4", channels)

    conv("conv1", M_CHANNELS, 1, 3)
    bn("bn1", M_CHANNELS)
    for name, in_planes, planes, stride, is_aff in block_specs():
        width = int(math.floor(planes * (BASE_WIDTH / 64.0)))
        conv(name + ".conv1", width * SCALE, in_planes, 1)
        bn(name + ".bn1", width * SCALE)
        for i in range(SCALE):
            conv(f"{name}.convs.{i}", width, width, 3)
            bn(f"{name}.bns.{i}", width)
        if is_aff:
            for i in range(SCALE - 1):
                aff(f"{name}.fuse_models.{i}", width)
        conv(name + ".conv3", planes * EXPANSION, width * SCALE, 1)
        bn(name + ".bn3", planes * EXPANSION)
        if stride != 1 or in_planes != planes * EXPANSION:
            conv(name + ".shortcut.0", planes * EXPANSION, in_planes, 1)
            bn(name + ".shortcut.1", planes * EXPANSION)
    conv("layer3_ds", M_CHANNELS * 8 * EXPANSION, M_CHANNELS * 4 * EXPANSION, 3)
    aff("fuse34", M_CHANNELS * 8 * EXPANSION)
    stats_dim = (FEAT_DIM // 8) * M_CHANNELS * 8 * EXPANSION * 2
    bound = 1.0 / math.sqrt(stats_dim)
    sd["seg_1.weight"] = (torch.rand(EMBED_DIM, stats_dim, generator=g) * 2 - 1) * bound
    sd["seg_1.bias"] = (torch.rand(EMBED_DIM, generator=g) * 2 - 1) * bound
    return sd



# ------------------------------------------------------------------------------------------------ Apollo restorer
def apollo_band_widths(sr=44100, win_ms=20):
    """Band split of look2hear/models/apollo.py:232-236: 79 bands of int(win / 160) bins + the rest."""
    win = int(sr * win_ms // 1000)
    enc_dim = win // 2 + 1
    bw = [int(win / 160)] * 79
    bw.append(enc_dim - sum(bw))
    return win, enc_dim, bw


def random_apollo_state_dict(seed=0, perturb=True, sr=44100, win_ms=20, feature_dim=256, layer=6):
    """Random-init weights with the key names / shapes of the reference's Apollo(sr, win, feature_dim, layer)
    (look2hear/models/apollo.py:215-257; checked key for key against the reference module by
    tests/test_apollo_oracle.py).  Conv1d default init; `perturb` randomises the RMSNorm gains."""
    g = torch.Generator().manual_seed(seed)
    _, _, bw = apollo_band_widths(sr, win_ms)
    N = feature_dim
    sd = {}

    def gain(name, n):
        sd[name] = 1.0 + 0.2 * torch.randn(n, generator=g) if perturb else torch.ones(n)

    def conv(name, out_c, in_c, k=1, bias=True, groups=1):
        fan = in_c // groups * k
        sd[name + ".weight"] = _uniform(g, (out_c, in_c // groups, k), fan)
        if bias:
            sd[name + ".bias"] = _uniform(g, (out_c,), fan)

    for i, w in enumerate(bw):
        gain(f"BN.{i}.0.weight", 2 * w + 1)
        conv(f"BN.{i}.1", N, 2 * w + 1)
    hd = N // 8
    freq = 1.0 / (10000 ** (torch.arange(0, hd, 2)[: hd // 2] / hd))
    ang = torch.arange(0, 100).reshape(-1, 1) * freq.reshape(1, -1)
    cos = torch.stack([torch.cos(ang)] * 2, -1).reshape(100, hd)
    sin = torch.stack([torch.sin(ang)] * 2, -1).reshape(100, hd)
    for l in range(layer):
        p = f"net.{l}.band_net."
        sd[p + "cos_freq"] = cos.clone()
        sd[p + "sin_freq"] = sin.clone()
        gain(p + "input_norm.weight", N)
        conv(p + "weight", 3 * N, N, bias=False)
        conv(p + "output", N, N, bias=False)
        gain(p + "MLP.0.weight", N)
        conv(p + "MLP.1", 8 * N, N, bias=False)
        conv(p + "MLP_output", N, 4 * N, bias=False)
        for b in range(3):
            q = f"net.{l}.seq_net.blocks.{b}.conv."
            conv(q + "0", N, N, k=7, groups=N)
            gain(q + "1.weight", N)
            conv(q + "2", 4 * N, N)
            conv(q + "4", N, 4 * N)
    for i, w in enumerate(bw):
        gain(f"output.{i}.0.weight", N)
        conv(f"output.{i}.1", 4 * w, N)
    return sd


def synthetic_fullband(n_items, n_samples, seed=4321, sr=44100):
    """A 44.1 kHz test signal for the restorer: a few decaying harmonics + low-passed noise, peak 0.5."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float32) / sr
    out = torch.zeros(n_items, n_samples)
    for i in range(n_items):
        f0 = 110.0 + 60.0 * float(torch.rand(1, generator=g))
        for h in range(1, 9):
            out[i] += torch.sin(2 * math.pi * f0 * h * t + 6.28 * float(torch.rand(1, generator=g))) / h
    x = torch.randn(n_items, n_samples, generator=g)
    X = torch.fft.rfft(x, dim=-1)
    f = torch.fft.rfftfreq(n_samples, 1.0 / sr)
    x = torch.fft.irfft(X / (1.0 + (f / 3000.0) ** 2), n=n_samples, dim=-1)
    out = out / out.abs().amax(dim=-1, keepdim=True) + 0.5 * x / x.abs().amax(dim=-1, keepdim=True)
    return (out / out.abs().amax(dim=-1, keepdim=True).clamp(min=1e-9) * 0.5).contiguous()
