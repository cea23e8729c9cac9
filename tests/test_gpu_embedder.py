"""GPU parity tests of the speaker-scoring path (fbank + ERes2NetV2 + cosine), through the C-ABI (libtdz.so),
against the oracle (torchaudio Kaldi fbank + oracle/eres2netv2_port.py, fp32 on the CPU).

ERes2NetV2 parity is against our own PyTorch restatement of the published architecture (the model lives in the
un-pinned, absent `modelscope` dependency: SURVEY.md section 8c)."""
import ctypes
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()))


def _snr(ref, est):
    import math
    ref, est = ref.double(), est.double()
    den = float(((ref - est) ** 2).sum())
    return 200.0 if den == 0 else 10 * math.log10(float((ref ** 2).sum()) / den)


def _record(name, value):
    path = os.path.join(ROOT, "gpurun_out", "parity.jsonl")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


@pytest.fixture(scope="module")
def setup():
    import torch
    from oracle import eres2netv2_port as E
    from targetdiarization_b200.synth import synthetic_mixture
    from targetdiarization_b200.embedder import Embedder
    sd = E.random_state_dict(seed=0)
    emb = Embedder(sd, "cuda:0")
    wav = synthetic_mixture(3, 16000 + 37, seed=5)  # 1.0023 s -> 98 frames (ragged vs the 160-sample hop)
    return torch, E, sd, emb, wav


def test_fbank_matches_kaldi(setup):
    torch, E, sd, emb, wav = setup
    ref = E.fbank_features(wav)
    out = emb.fbank(wav.cuda()).cpu()
    assert out.shape == ref.shape
    err = float((out - ref).abs().max())
    snr = _snr(ref, out)
    _record("fbank_snr_db", snr)
    _record("fbank_max_abs_err", err)
    # log-mel features, fp32 FFT: tolerance 1e-3 absolute on values of magnitude ~1..10
    assert err < 1e-3 and snr > 70, (err, snr)


@pytest.mark.parametrize("stop", [-1, 0, 2, 3, 6, 7, 12, 13, 15, 16])
def test_block_parity(setup, stop):
    """Feature map after the stem / a residual block / fuse34 versus the oracle's taps (bf16 tensor-core operands,
    fp32 accumulate and residual stream: tolerance 35 dB per map)."""
    torch, E, sd, emb, wav = setup
    feat = E.fbank_features(wav)
    taps = {}
    with torch.no_grad():
        E.eres2netv2_forward(sd, feat, taps)
    names = ["stem"] + [s[0] for s in E.block_specs()] + ["fuse34"]
    ref = taps[names[stop + 1]].permute(0, 2, 3, 1).contiguous()  # NCHW -> NHWC
    # feature maps are stored in bf16 on the device (the fuse34 map, which feeds the pooling, in fp32)
    out = torch.empty(ref.shape, dtype=torch.float32 if stop == 16 else torch.bfloat16, device="cuda")
    N, frames, _ = feat.shape
    lib, h = emb._h.lib, emb._h
    nbytes = int(lib.tdz_embed_workspace_bytes(N, frames))
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device="cuda")
    raw.fill_(0xFF)
    ws = raw[(-raw.data_ptr()) % 1024:][:nbytes]   # the C ABI wants a 1024 B aligned workspace
    f = feat.cuda().contiguous()
    h.check(lib.tdz_embed_debug(h.ptr, f.data_ptr(), N, frames, out.data_ptr(), ws.data_ptr(), nbytes,
                                torch.cuda.current_stream().cuda_stream, stop), "tdz_embed_debug")
    torch.cuda.synchronize()
    snr = _snr(ref, out.float().cpu())
    _record(f"sv_block_{names[stop + 1]}_snr_db", snr)
    assert snr > 35, snr


def test_embedding_cosine(setup):
    """north_star tolerance: embedding cosine >= 0.999 versus the reference implementation."""
    torch, E, sd, emb, wav = setup
    ref = E.embed(sd, wav)
    out = emb.embed_many(wav.cuda()).cpu()
    assert out.shape == (3, 192)
    cs = [_cos(ref[i], out[i]) for i in range(3)]
    _record("embedding_cosine_min", min(cs))
    assert min(cs) >= 0.999, cs


def test_reference_call_surface(setup):
    """`self.embedding['eres2netv2_large'](wav.reshape(1,-1), output_emb=True)['embs'].reshape(-1)`
    (TargetASR.py:155-163) and the int16 convention."""
    torch, E, sd, emb, wav = setup
    x = wav[0].numpy()
    r = emb(x.reshape(1, -1), output_emb=True)
    assert r["embs"].shape == (1, 192) and r["embs"].dtype == np.float32
    ref = E.embed(sd, wav[:1])[0]
    assert _cos(ref, torch.from_numpy(r["embs"][0])) >= 0.999
    # ragged list input, grouped by length
    many = emb.embed_many([wav[0, :12000], wav[1], wav[2, :12000]]).cpu()
    ref2 = E.embed(sd, wav[[0, 2], :12000])
    assert _cos(ref2[0], many[0]) >= 0.999 and _cos(ref2[1], many[2]) >= 0.999


def test_cosine_scores_semantics(setup):
    """TargetASR.cosine_similarity: zero vector -> 1.0, clamp to [0,1] (TargetASR.py:144-152)."""
    torch, E, sd, emb, wav = setup
    from oracle.stage_port import cosine_similarity
    g = torch.Generator().manual_seed(3)
    e = torch.randn(6, 192, generator=g)
    e[2] = 0.0
    tgt = torch.randn(192, generator=g)
    e[4] = -tgt  # negative similarity -> clamped to 0
    e[5] = tgt * 3
    got = emb.cosine_scores(e.cuda(), tgt.cuda()).cpu()
    want = [cosine_similarity(e[i].numpy(), tgt.numpy()) for i in range(6)]
    assert got[2] == 1.0 and got[4] == 0.0
    assert np.allclose(got.numpy(), np.array(want, dtype=np.float32), atol=1e-6)
    z = emb.cosine_scores(e.cuda(), torch.zeros(192).cuda()).cpu()
    assert bool((z == 1.0).all())


def test_too_short_raises(setup):
    torch, E, sd, emb, wav = setup
    with pytest.raises(ValueError):
        emb.embed_many(torch.zeros(1, 300).cuda())
