"""CPU tests of the SURVEY.md 8f-4 oracle (oracle/apollo_port.py): the Apollo restatement and the MDX STFT / iSTFT
restatement against golden vectors produced by RUNNING THE REFERENCE (tests/golden/apollo_small.npz, mdx_stft.npz,
oracle/make_golden.py), against the reference module itself where the tree exists, and the host-side checks of
`Restorer` / `ConvTDFNet` that need no GPU."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import needs_reference
from oracle import apollo_port as AP
from targetdiarization_b200 import synth
from targetdiarization_b200.restorer import check_apollo_args
from targetdiarization_b200.weights import apollo_key_shapes, check_apollo_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sd():
    return synth.random_apollo_state_dict(0)


def test_port_matches_reference_golden(sd):
    g = np.load(os.path.join(GOLDEN, "apollo_small.npz"))
    xa = synth.synthetic_fullband(2, 22050 + 123, seed=4321).reshape(1, 2, -1)
    xb = synth.synthetic_fullband(2, 4500, seed=77).reshape(2, 1, -1)
    with torch.no_grad():
        ya, yb = AP.apollo_forward(sd, xa), AP.apollo_forward(sd, xb)
    # fp32 restatement in another summation order: >= 100 dB against the reference module's output
    assert AP.snr_db(torch.from_numpy(g["out_a"]), ya) > 100
    assert AP.snr_db(torch.from_numpy(g["out_b"]), yb) > 100


def test_bf16_operand_emulation_meets_the_contract(sd):
    """The CUDA path feeds bf16 operands to the tensor cores; emulated on the CPU this stays above the 40 dB bar."""
    g = np.load(os.path.join(GOLDEN, "apollo_small.npz"))
    xb = synth.synthetic_fullband(2, 4500, seed=77).reshape(2, 1, -1)
    with torch.no_grad():
        yb = AP.apollo_forward(sd, xb, operands="bf16")
    assert AP.snr_db(torch.from_numpy(g["out_b"]), yb) > 43


def test_mdx_stft_port_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "mdx_stft.npz"))
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 2, int(g["chunk_size"]), generator=gen) * 0.1
    spec_in = torch.randn(2, 4, 3072, 8, generator=gen)
    assert np.array_equal(x.numpy(), g["x"])
    spec = AP.mdx_stft(x, 6144, 1024, 3072)
    assert spec.shape == (2, 4, 3072, 8)
    assert np.array_equal(spec[:, :, ::5].numpy(), g["spec_bins5"])   # the same torch.stft call: bit-identical
    wav = AP.mdx_istft(spec_in, 6144, 1024)
    assert np.array_equal(wav.numpy(), g["wav"])


def test_key_table_and_arg_checks(sd):
    want = apollo_key_shapes()
    assert set(want) == set(sd) and all(tuple(sd[k].shape) == want[k] for k in want)
    check_apollo_state_dict(sd)
    bad = dict(sd)
    bad.pop("net.3.band_net.output.weight")
    with pytest.raises(RuntimeError):
        check_apollo_state_dict(bad)
    bad = dict(sd)
    bad["output.79.1.weight"] = torch.zeros(20, 256, 1)
    with pytest.raises(RuntimeError):
        check_apollo_state_dict(bad)
    check_apollo_args(sr=44100, win=20, feature_dim=256, layer=6)     # AudioProcessor.py:279
    check_apollo_args(44100, 20)
    with pytest.raises(ValueError):
        check_apollo_args(sr=44100, win=20, feature_dim=128, layer=6)
    with pytest.raises(ValueError):
        check_apollo_args(sr=16000)
    with pytest.raises(TypeError):
        check_apollo_args(depth=6)


@needs_reference
def test_port_and_key_table_against_reference_module(sd):
    from oracle import ref_loader
    m = ref_loader.build_reference_apollo()
    ref_sd = m.state_dict()
    want = apollo_key_shapes()
    assert set(ref_sd) == set(want)
    assert all(tuple(v.shape) == want[k] for k, v in ref_sd.items())
    for k in ref_sd:   # the generator reproduces the module's rotary buffers
        if k.endswith("_freq"):
            assert torch.allclose(sd[k], ref_sd[k], atol=1e-6)
    m.load_state_dict(sd)
    x = synth.synthetic_fullband(3, 3000, seed=9).reshape(1, 3, -1)
    with torch.no_grad():
        assert AP.snr_db(m(x), AP.apollo_forward(sd, x)) > 100


@needs_reference
def test_mdx_port_against_reference_class():
    from oracle import ref_loader
    net = ref_loader.load_reference_conv_tdf_net()(target_name="vocals", L=11, dim_f=2048, dim_t=4, n_fft=4096,
                                                   hop=512, device="cpu")
    x = torch.randn(3, 2, net.chunk_size)
    assert torch.equal(net.stft(x), AP.mdx_stft(x, 4096, 512, 2048))
    s = torch.randn(3, 4, 2048, 16)
    assert torch.equal(net.istft(s), AP.mdx_istft(s, 4096, 512))
