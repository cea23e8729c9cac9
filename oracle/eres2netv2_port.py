"""PyTorch restatement of the ERes2NetV2-Large speaker embedder + its Kaldi-fbank front-end.

TEST INFRASTRUCTURE ONLY (oracle).  PARITY UNPINNED: the model lives in the un-pinned third-party dependency
`modelscope` (requirements.txt:25 of the reference; pipeline id iic/speech_eres2netv2w24s4ep4_sv_zh-cn_16k-common,
README.md:144), which is absent from /root/reference and from this image, and the reference holds no golden
vector for it.  This file restates the published architecture (3D-Speaker / modelscope `ERes2NetV2`,
feat_dim=80, embedding_size=192, m_channels=64, num_blocks=[3,4,6,3], baseWidth=24, scale=4, expansion=4,
TSTP pooling, two_emb_layer=False) and is anchored on the reference's call site contract
(TargetASR.py:102-103,155-163: ndarray [1,T] -> {'embs': float32 [1,192]}).  The open points listed in
SURVEY.md section 8c (TSTP epsilon, AFF ordering, Hardtanh(0,20), stride placement) are *defined by this
restatement*.  The fbank oracle is torchaudio.compliance.kaldi.fbank (in the image) with the pipeline's defaults.

State-dict key names follow the published module names (conv1, bn1, layer1.0.conv1, layer3.0.fuse_models.0.local_att.0,
layer3_ds, fuse34.local_att.*, seg_1) so that a real checkpoint would load unchanged.
"""
import math

import torch
import torch.nn.functional as F

NUM_BLOCKS = (3, 4, 6, 3)
M_CHANNELS = 64
BASE_WIDTH = 24
SCALE = 4
EXPANSION = 4
FEAT_DIM = 80
EMBED_DIM = 192
BN_EPS = 1e-5


def fbank_features(wav):
    """wav float32 [N,T] in [-1,1] -> [N, frames, 80]: Kaldi fbank(80) + per-utterance mean normalisation."""
    import torchaudio.compliance.kaldi as kaldi
    feats = []
    for w in wav:
        f = kaldi.fbank(w.unsqueeze(0), num_mel_bins=FEAT_DIM)
        feats.append(f - f.mean(dim=0, keepdim=True))
    return torch.stack(feats)


# The oracle's OWN architecture table, written out row by row from the published ERes2NetV2 definition
# (m_channels 64, num_blocks [3,4,6,3], expansion 4; stride 2 on the first block of layers 2-4; AFF joins in layers
# 3-4) - deliberately not imported from the product package, so that a wrong table on either side shows up as a
# parity failure (tests/test_oracle_port.py also checks the weight generator against THIS table).
#   (name, in_planes, planes, stride, joins sub-bands by AFF)
BLOCK_SPECS = (
    ("layer1.0", 64, 64, 1, False), ("layer1.1", 256, 64, 1, False), ("layer1.2", 256, 64, 1, False),
    ("layer2.0", 256, 128, 2, False), ("layer2.1", 512, 128, 1, False), ("layer2.2", 512, 128, 1, False),
    ("layer2.3", 512, 128, 1, False),
    ("layer3.0", 512, 256, 2, True), ("layer3.1", 1024, 256, 1, True), ("layer3.2", 1024, 256, 1, True),
    ("layer3.3", 1024, 256, 1, True), ("layer3.4", 1024, 256, 1, True), ("layer3.5", 1024, 256, 1, True),
    ("layer4.0", 1024, 512, 2, True), ("layer4.1", 2048, 512, 1, True), ("layer4.2", 2048, 512, 1, True),
)


def block_specs():
    return list(BLOCK_SPECS)


def expected_key_shapes():
    """{key: shape} of the state dict this restatement consumes, derived from BLOCK_SPECS alone."""
    sh = {"conv1.weight": (64, 1, 3, 3)}

    def bn(name, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            sh[f"{name}.{k}"] = (c,)

    def aff(name, c):
        sh[name + ".local_att.0.weight"] = (c // 4, 2 * c, 1, 1)
        sh[name + ".local_att.0.bias"] = (c // 4,)
        bn(name + ".local_att.1", c // 4)
        sh[name + ".local_att.3.weight"] = (c, c // 4, 1, 1)
        sh[name + ".local_att.3.bias"] = (c,)
        bn(name + ".local_att.4", c)

    bn("bn1", 64)
    for name, in_planes, planes, stride, is_aff in BLOCK_SPECS:
        width = planes * BASE_WIDTH // 64
        sh[name + ".conv1.weight"] = (width * SCALE, in_planes, 1, 1)
        bn(name + ".bn1", width * SCALE)
        for i in range(SCALE):
            sh[f"{name}.convs.{i}.weight"] = (width, width, 3, 3)
            bn(f"{name}.bns.{i}", width)
        if is_aff:
            for i in range(SCALE - 1):
                aff(f"{name}.fuse_models.{i}", width)
        sh[name + ".conv3.weight"] = (planes * EXPANSION, width * SCALE, 1, 1)
        bn(name + ".bn3", planes * EXPANSION)
        if stride != 1 or in_planes != planes * EXPANSION:
            sh[name + ".shortcut.0.weight"] = (planes * EXPANSION, in_planes, 1, 1)
            bn(name + ".shortcut.1", planes * EXPANSION)
    sh["layer3_ds.weight"] = (2048, 1024, 3, 3)
    aff("fuse34", 2048)
    sh["seg_1.weight"] = (EMBED_DIM, 2 * 2048 * (FEAT_DIM // 8))
    sh["seg_1.bias"] = (EMBED_DIM,)
    return sh


def random_state_dict(seed=0):
    """The shared synthetic-weight generator (data, not arithmetic); its keys / shapes are checked against
    expected_key_shapes() above before use."""
    from targetdiarization_b200.synth import random_eres2netv2_state_dict
    sd = random_eres2netv2_state_dict(seed=seed)
    want = expected_key_shapes()
    got = {k: tuple(v.shape) for k, v in sd.items()}
    if got != want:
        diff = sorted(set(got.items()) ^ set(want.items()))[:6]
        raise AssertionError(f"weight generator and oracle architecture table disagree: {diff}")
    return sd


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, BN_EPS)


def _relu20(x):
    return F.hardtanh(x, 0.0, 20.0)


def _aff(x, y, sd, name):
    """AFF: att = 1 + tanh(BN(Conv1x1(SiLU(BN(Conv1x1(cat(x,y))))))); x*att + y*(2-att)."""
    xa = torch.cat((x, y), dim=1)
    a = F.conv2d(xa, sd[name + ".local_att.0.weight"], sd[name + ".local_att.0.bias"])
    a = F.silu(_bn(a, sd, name + ".local_att.1"))
    a = F.conv2d(a, sd[name + ".local_att.3.weight"], sd[name + ".local_att.3.bias"])
    a = 1.0 + torch.tanh(_bn(a, sd, name + ".local_att.4"))
    return x * a + y * (2.0 - a)


def _block(x, sd, name, in_planes, planes, stride, is_aff):
    width = int(math.floor(planes * (BASE_WIDTH / 64.0)))
    out = _relu20(_bn(F.conv2d(x, sd[name + ".conv1.weight"], stride=stride), sd, name + ".bn1"))
    spx = torch.split(out, width, 1)
    outs = []
    sp = None
    for i in range(SCALE):
        if i == 0:
            sp = spx[i]
        elif is_aff:
            sp = _aff(sp, spx[i], sd, f"{name}.fuse_models.{i - 1}")
        else:
            sp = sp + spx[i]
        sp = F.conv2d(sp, sd[f"{name}.convs.{i}.weight"], padding=1)
        sp = _relu20(_bn(sp, sd, f"{name}.bns.{i}"))
        outs.append(sp)
    out = torch.cat(outs, 1)
    out = _bn(F.conv2d(out, sd[name + ".conv3.weight"]), sd, name + ".bn3")
    if (name + ".shortcut.0.weight") in sd:
        res = _bn(F.conv2d(x, sd[name + ".shortcut.0.weight"], stride=stride), sd, name + ".shortcut.1")
    else:
        res = x
    return _relu20(out + res)


def eres2netv2_forward(sd, feats, taps=None):
    """feats [N, frames, 80] -> embeddings [N, 192]."""
    x = feats.permute(0, 2, 1).unsqueeze(1)  # [N,1,80,frames]
    out = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], padding=1), sd, "bn1"))
    if taps is not None:
        taps["stem"] = out
    layer_out = {}
    for name, in_planes, planes, stride, is_aff in block_specs():
        out = _block(out, sd, name, in_planes, planes, stride, is_aff)
        layer_out[name.split(".")[0]] = out
        if taps is not None:
            taps[name] = out
    out3, out4 = layer_out["layer3"], layer_out["layer4"]
    out3_ds = F.conv2d(out3, sd["layer3_ds.weight"], stride=2, padding=1)
    fuse = _aff(out4, out3_ds, sd, "fuse34")
    if taps is not None:
        taps["fuse34"] = fuse
    n = fuse.shape[0]
    z = fuse.reshape(n, -1, fuse.shape[-1])  # [N, C*F, frames/8]
    mean = z.mean(dim=-1)
    std = torch.sqrt(z.var(dim=-1, unbiased=True) + 1e-8)
    stats = torch.cat((mean, std), dim=-1)
    if taps is not None:
        taps["stats"] = stats
    return F.linear(stats, sd["seg_1.weight"], sd["seg_1.bias"])


def embed(sd, wav):
    """wav float32 [N,T] -> [N,192] (fbank + model), the oracle of Embedder.embed_many."""
    with torch.no_grad():
        return eres2netv2_forward(sd, fbank_features(wav))


def cosine_similarity(a, b):
    """TargetASR.cosine_similarity (TargetASR.py:144-152) restated with numpy semantics."""
    import numpy as np
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if np.all(a == 0.0) or np.all(b == 0.0):
        return 1.0
    s = float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))
    return float(max(0.0, min(s, 1.0)))


def pick_target(spk1_score, spk2_score, threshold=0.0):
    """multi_speakers_separate_asr core (TargetASR.py:612-625): None if both below threshold, else
    1 iff spk1_score > spk2_score (strict), else 2."""
    if spk1_score < threshold and spk2_score < threshold:
        return None
    return 1 if spk1_score > spk2_score else 2
