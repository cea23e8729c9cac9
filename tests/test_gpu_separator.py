"""GPU parity tests of the separator, through the C-ABI (libtdz.so) against the oracle port."""
import os

import pytest

from tests import sep_steps

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def step_results():
    out = os.path.join(ROOT, "gpurun_out", "steps.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    return sep_steps.run_all(out)


@pytest.mark.parametrize("step", sep_steps.STEPS)
def test_step_parity(step_results, step):
    """Each launch step of tdz_separate alone, oracle inputs -> oracle outputs (layer 0, B=2, T=9613)."""
    r = step_results[step]
    assert r["ok"], f"{step}: {r.get('metrics')} {r.get('error', '')}"


def _full(B, T, seed, min_db=40.0):
    import torch
    from oracle.mossformer2_port import mossformer2_forward, snr_db
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    sd = random_state_dict(seed=seed)
    g = torch.Generator().manual_seed(1234 + seed)
    mix = torch.randn(B, T, generator=g) * 0.1
    sep = Separator(sd, "cuda:0")
    out = sep(mix.cuda()).cpu()
    with torch.no_grad():
        ref = mossformer2_forward(sd, mix)
    snr = snr_db(ref, out)
    print(f"full forward B={B} T={T}: {snr:.2f} dB")
    import json
    with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": f"separator_full_B{B}_T{T}_seed{seed}_snr_db", "value": snr}) + "\n")
    assert out.shape == (B, 2, T)
    # north_star tolerance: waveform SNR >= 40 dB versus the reference output
    assert snr >= min_db, f"{snr:.2f} dB"
    return out


def test_full_forward_small():
    _full(1, 8000, 0)


def test_full_forward_production_window():
    """The window AudioProcessor.separate_speaker actually feeds (160 000 samples = 10 s, S = 19 999 frames, 79 attention
    groups, 10 fixed-size linear-attention splits): the precision budget is tightest here."""
    _full(1, 160000, 3)


def test_full_forward_batch_ragged_tail():
    # T = 9613 -> S = 1200, T' = 9608: the last 5 samples are the zero pad of mossformer2.py:585-586
    out = _full(2, 9613, 1)
    assert float(out[..., 9608:].abs().max()) == 0.0


@pytest.mark.parametrize("T", [6000, 140000])
def test_batch_equals_singles(T):
    """Chunks are independent units: a batch of 2 equals two single calls bit for bit (T = 140 000 spans several
    fixed-size splits of the linear-attention sum, whose order must not depend on the batch)."""
    import torch
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    sd = random_state_dict(seed=2)
    g = torch.Generator().manual_seed(7)
    mix = (torch.randn(2, T, generator=g) * 0.1).cuda()
    sep = Separator(sd, "cuda:0")
    both = sep(mix).clone()
    one0 = sep(mix[0:1]).clone()
    one1 = sep(mix[1:2]).clone()
    assert torch.equal(both[0], one0[0]) and torch.equal(both[1], one1[0])


def test_real_speech_excerpt_against_reference_run():
    """CUDA path vs the output of the reference module itself (tests/golden/chat_mix_excerpt.npz, generated in the
    build container from assets/chat_mix.wav, config C1): waveform SNR >= 40 dB."""
    import json
    import numpy as np
    import torch
    from oracle.mossformer2_port import snr_db
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    gd = np.load(os.path.join(ROOT, "tests", "golden", "chat_mix_excerpt.npz"))
    sep = Separator(random_state_dict(seed=0, perturb=True), "cuda:0")
    x = torch.from_numpy(gd["pcm"].astype(np.float32) / 32768.0)[None]
    y = sep(x.cuda()).cpu()
    snr = snr_db(torch.from_numpy(gd["out"]), y)
    with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": "separator_chat_mix_excerpt_vs_reference_run_snr_db", "value": snr}) + "\n")
    assert snr >= 40.0, snr


@pytest.mark.parametrize("T", [100, 2055, 6400])
def test_short_inputs(T):
    """Short inputs under the 40 dB contract: 12 frames (T = 100), just over one attention group (T = 2 055) and the
    shortest clip the reference's own chunk loop can process (0.4 s = 6 400 samples: pyloudnorm's minimum,
    AudioProcessor.py:949-950)."""
    _full(1, T, 4, min_db=40.0)


@pytest.mark.parametrize("T", [16, 23])
def test_one_frame_inputs_are_finite(T):
    """T = 16 .. 23 is ONE encoder frame: InstanceNorm over time then divides by sqrt(0 + eps) and GroupNorm sees 512
    values, so the problem is ill-conditioned - the reference's own fp32 result moves by several dB with the summation
    order.  Scanned on B200 over 3 weight seeds x 4 inputs (tools/tiny_snr.py): 31 .. 54 dB for T <= 23, 39.8 .. 47.9 dB
    for T = 40, >= 40.9 dB from T = 100 on.  Not reachable through separate_speaker (0.4 s minimum); checked here as a
    robustness case: finite output of the right shape and a 30 dB floor."""
    out = _full(1, T, 4, min_db=30.0)
    import torch
    assert bool(torch.isfinite(out).all())


def test_too_short_input_is_an_error():
    import torch
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    sep = Separator(random_state_dict(seed=0), "cuda:0")
    with pytest.raises(RuntimeError):
        sep(torch.zeros(1, 15).cuda())


@pytest.mark.parametrize("B,T", [(1, 9613), (3, 40000)])
def test_back_to_back_linear_project_equals_two_kernels(B, T):
    """fsmn.linear -> ReLU -> fsmn.project (fsmn.py:131-139) runs as ONE back-to-back GEMM in the forward (hidden
    activations stay in TMEM / shared memory); asked for one step at a time the library runs the two-kernel form.
    Same k order, same bf16 rounding of the hidden activations: the `p` buffer must agree bit for bit."""
    import torch
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    sep = Separator(random_state_dict(seed=3), "cuda:0")
    mix = (torch.randn(B, T, generator=torch.Generator().manual_seed(1)) * 0.1).cuda()
    sep(mix)                                   # populates the workspace (xubf of the last layer)
    k = sep.STEP_NAMES.index("FSMN_LIN")
    S = sep.layout(B, T).S
    sep.debug_buffer(B, T, "p", torch.float32, 256).zero_()
    sep(mix, _debug=(1, k, k + 1))             # fused
    fused = sep.debug_buffer(B, T, "p", torch.float32, 256)[:, :S].clone()
    sep.debug_buffer(B, T, "p", torch.float32, 256).zero_()
    sep(mix, _debug=(1, k, k))
    sep(mix, _debug=(1, k + 1, k + 1))         # linear, then project
    two = sep.debug_buffer(B, T, "p", torch.float32, 256)[:, :S].clone()
    assert float(two.abs().max()) > 0
    assert torch.equal(fused, two)
    # the cta_group::2 form of the same kernel (a CTA pair per 256 rows; opt-in, read when the handle is created)
    import os
    os.environ["TDZ_B2B_CG2"] = "1"
    try:
        sep2 = Separator(random_state_dict(seed=3), "cuda:0")
    finally:
        del os.environ["TDZ_B2B_CG2"]
    sep2(mix)
    sep2.debug_buffer(B, T, "p", torch.float32, 256).zero_()
    sep2(mix, _debug=(1, k, k + 1))
    assert torch.equal(sep2.debug_buffer(B, T, "p", torch.float32, 256)[:, :S], two)
