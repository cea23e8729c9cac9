"""Build-container-only tests (skipped where /root/reference is absent): the oracle restatements against the
reference's own executable modules on fresh seeds, and the synthetic state dict against the reference's keys."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, stage_port
from oracle.mossformer2_port import mossformer2_forward, snr_db
from targetdiarization_b200 import synth
from tests.conftest import needs_reference

pytestmark = needs_reference


@pytest.fixture(scope="module")
def ref_model():
    return ref_loader.build_reference_mossformer2(seed=0)


def test_synth_state_dict_has_reference_keys_and_shapes(ref_model):
    ref = ref_model.state_dict()
    ours = synth.random_state_dict(seed=3)
    assert len(ref) == 1099
    assert set(ref.keys()) == set(ours.keys())
    for k, v in ref.items():
        assert tuple(v.shape) == tuple(ours[k].shape), k


def test_port_matches_reference_default_init(ref_model):
    """Reference default initialisation (torch.manual_seed(0); MossFormer2()), its own state dict through the port."""
    g = torch.Generator().manual_seed(11)
    mix = torch.randn(2, 3001, generator=g) * 0.1
    with torch.no_grad():
        want = ref_model(mix)
        got = mossformer2_forward(ref_model.state_dict(), mix)
    assert want.shape == got.shape == (2, 2, 3001)
    assert snr_db(want, got) >= 100.0


def test_port_accepts_the_three_input_ranks(ref_model):
    sd = ref_model.state_dict()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(1600, generator=g) * 0.1
    with torch.no_grad():
        a = mossformer2_forward(sd, x)
        b = mossformer2_forward(sd, x[None])
        c = mossformer2_forward(sd, x[None, None])
        want = ref_model(x)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert snr_db(want, a) >= 100.0


def test_prefix_is_not_the_full_chunk(ref_model):
    """Negative control: the output of a chunk depends on the whole chunk, so chunk
    boundaries are part of the numerical contract."""
    sd = ref_model.state_dict()
    g = torch.Generator().manual_seed(13)
    x = torch.randn(1, 4000, generator=g) * 0.1
    with torch.no_grad():
        full = mossformer2_forward(sd, x)
        half = mossformer2_forward(sd, x[:, :2000])
    assert snr_db(full[..., :2000], half) < 30.0


def test_wav_chunk_inference_port_vs_reference_function():
    wci = ref_loader.load_reference_wav_chunk_inference()
    g = torch.Generator().manual_seed(5)
    for L in (100, 3999, 4000, 12000, 12001, 25777):
        mix = torch.randn(1, 1, L, generator=g)

        def model(x):
            return torch.stack((x * 0.5, torch.tanh(x)), dim=1)
        want = wci(model, mix, sr=1000, n_tracks=2)
        got = stage_port.wav_chunk_inference(model, mix, sr=1000)
        assert torch.equal(want, got), L


def test_chunk_rule_vs_reference_function():
    from oracle.make_golden import reference_separate_speaker
    run = reference_separate_speaker()
    rng = np.random.default_rng(0)
    lengths = [int(v) for v in rng.integers(1, 900000, size=12)] + [160000 * k + d for k in (1, 2, 3) for d in (-1, 0, 1)]
    for L in lengths:
        _, _, b = run(np.zeros(L, dtype=np.float32), lambda x: torch.zeros(1, 2, x.shape[-1]), lambda a: 0.0)
        assert b == stage_port.chunk_bounds(L), L
