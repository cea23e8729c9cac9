"""GPU parity tests of SURVEY.md 8f-4 through the C-ABI (libtdz.so): the STFT / inverse STFT engine behind
`ConvTDFNet` (AudioProcessor.py:65-120) and the Apollo restorer (look2hear/models/apollo.py) behind `Restorer`, against
oracle/apollo_port.py on the CPU and against golden vectors produced by running the reference
(tests/golden/apollo_small.npz, mdx_stft.npz).

Tolerances: the transforms are fp32 (no reduced-precision operand) - >= 100 dB; the Apollo forward feeds bf16 operands
to the tensor cores like the separator - the north_star's waveform bar, SNR >= 40 dB."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _record(name, value):
    path = os.path.join(ROOT, "gpurun_out", "parity.jsonl")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


@pytest.fixture(scope="module")
def env():
    import torch
    from oracle import apollo_port as AP
    from targetdiarization_b200 import ConvTDFNet, Restorer, synth
    sd = synth.random_apollo_state_dict(0)
    return torch, AP, synth, sd, Restorer(sd, "cuda:0"), ConvTDFNet


# ------------------------------------------------------------------------------------------------ STFT engine
@pytest.mark.parametrize("n_fft,hop,dim_f,dim_t", [(6144, 1024, 3072, 3), (6144, 1024, 3072, 8), (4096, 256, 2048, 5),
                                                    (7680, 2048, 3000, 4)])
def test_mdx_stft_istft_match_oracle(env, n_fft, hop, dim_f, dim_t):
    torch, AP, synth, sd, rest, ConvTDFNet = env
    net = ConvTDFNet("vocals", 11, dim_f, dim_t, n_fft, hop, "cuda:0")
    B = 1 if dim_t == 8 else 3
    g = torch.Generator().manual_seed(n_fft + dim_t)
    x = torch.randn(B, 2, net.chunk_size, generator=g) * 0.1
    ref = AP.mdx_stft(x, n_fft, hop, dim_f)
    out = net.stft(x.cuda()).cpu()
    assert out.shape == ref.shape == (B, 4, dim_f, 2 ** dim_t)
    snr = AP.snr_db(ref, out)
    _record(f"mdx_stft_snr_db[{n_fft},{hop},{dim_t}]", snr)
    assert snr > 100, snr
    s = torch.randn(B, 4, dim_f, 2 ** dim_t, generator=g)
    ref_w = AP.mdx_istft(s, n_fft, hop)
    out_w = net.istft(s.cuda())
    assert out_w.shape == ref_w.shape == (B, 2, net.chunk_size) and not out_w.is_cuda
    snr = AP.snr_db(ref_w, out_w)
    _record(f"mdx_istft_snr_db[{n_fft},{hop},{dim_t}]", snr)
    assert snr > 100, snr


def test_mdx_golden_from_reference_class(env):
    torch, AP, synth, sd, rest, ConvTDFNet = env
    g = np.load(os.path.join(GOLDEN, "mdx_stft.npz"))
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 2, int(g["chunk_size"]), generator=gen) * 0.1
    spec_in = torch.randn(2, 4, 3072, 8, generator=gen)
    net = ConvTDFNet("vocals", 11, 3072, 3, 6144, 1024, "cuda:0")
    spec = net.stft(x.cuda()).cpu()
    snr = AP.snr_db(torch.from_numpy(g["spec_bins5"]), spec[:, :, ::5])
    _record("mdx_stft_vs_reference_class_snr_db", snr)
    assert snr > 100, snr
    snr = AP.snr_db(torch.from_numpy(g["wav"]), net.istft(spec_in.cuda()))
    _record("mdx_istft_vs_reference_class_snr_db", snr)
    assert snr > 100, snr


def test_mdx_round_trip_at_production_size(env):
    """denoise_vocal's shape (AudioProcessor.py:241: dim_t 2^8, hop 1024 -> 261 120-sample stereo chunks, 4 of them):
    istft(stft(x)) == x away from the chunk edges when every bin is kept (size-independent property, no oracle)."""
    torch, AP, synth, sd, rest, ConvTDFNet = env
    net = ConvTDFNet("vocals", 11, 3073, 8, 6144, 1024, "cuda:0")
    x = torch.randn(4, 2, net.chunk_size, generator=torch.Generator().manual_seed(3)) * 0.1
    y = net.istft(net.stft(x.cuda()))
    snr = AP.snr_db(x, y)
    _record("mdx_round_trip_snr_db", snr)
    assert snr > 100, snr


def test_apollo_stft_frame_major(env):
    torch, AP, synth, sd, rest, _ = env
    for ns in (22173, 4500, 882, 443):   # T odd / even, one window, the minimum the reflect padding allows + 1
        x = synth.synthetic_fullband(2, ns, seed=ns).reshape(1, 2, ns)
        ref = torch.view_as_real(AP.stft(x.reshape(2, ns), 882, 441))
        out = rest.tap(x.cuda(), "spec").cpu()
        assert out.shape == ref.shape
        snr = AP.snr_db(ref, out)
        _record(f"apollo_stft_snr_db[{ns}]", snr)
        assert snr > 100, (ns, snr)


# ------------------------------------------------------------------------------------------------ Apollo
def test_apollo_taps_match_oracle(env):
    """Every stage of the forward against the oracle's token-major taps (one run localises a divergence)."""
    torch, AP, synth, sd, rest, _ = env
    x = synth.synthetic_fullband(2, 4500, seed=77).reshape(2, 1, -1)
    taps = {}
    with torch.no_grad():
        AP.apollo_forward(sd, x, taps=taps)
    bars = {"feat": 90, "att0": 40, "band0": 45, "layer0": 43, "layer1": 42, "layer2": 41, "layer3": 40, "layer4": 40,
            "layer5": 40, "est_spec": 40}
    got = {}
    for name, bar in bars.items():
        ref = taps[name]
        if name == "est_spec":
            ref = torch.view_as_real(ref)
        if name == "att0":
            ref = ref.reshape(2, -1, 80, 256)
        out = rest.tap(x.cuda(), name).cpu()
        assert out.shape == ref.shape, (name, out.shape, ref.shape)
        got[name] = AP.snr_db(ref, out)
        _record(f"apollo_tap_snr_db[{name}]", got[name])
    assert all(got[n] > bars[n] for n in bars), got


def test_apollo_forward_matches_reference_golden(env):
    torch, AP, synth, sd, rest, _ = env
    g = np.load(os.path.join(GOLDEN, "apollo_small.npz"))
    xa = synth.synthetic_fullband(2, 22050 + 123, seed=4321).reshape(1, 2, -1)
    xb = synth.synthetic_fullband(2, 4500, seed=77).reshape(2, 1, -1)
    for name, x in (("a", xa), ("b", xb)):
        out = rest(x.cuda())
        assert out.is_cuda and out.shape == x.shape
        snr = AP.snr_db(torch.from_numpy(g["out_" + name]), out.cpu())
        _record(f"apollo_vs_reference_module_snr_db[{name}]", snr)
        assert snr >= 40, (name, snr)


def test_apollo_batch_equals_singles_and_longer_input(env):
    """Rows are independent: a batch is bit-identical to its rows run alone; a 6 s input (601 frames, 48 080 tokens)
    against the oracle."""
    torch, AP, synth, sd, rest, _ = env
    x = synth.synthetic_fullband(3, 9000, seed=11).reshape(3, 1, -1).cuda()
    full = rest(x)
    for i in range(3):
        assert torch.equal(full[i:i + 1], rest(x[i:i + 1]))
    xl = synth.synthetic_fullband(1, 6 * 44100, seed=12).reshape(1, 1, -1)
    with torch.no_grad():
        ref = AP.apollo_forward(sd, xl)
    snr = AP.snr_db(ref, rest(xl.cuda()).cpu())
    _record("apollo_6s_snr_db", snr)
    assert snr >= 40, snr


def test_apollo_chunked_equals_single_pass(env):
    """Bounded memory: with a workspace below the single-pass size the library runs the network over frame chunks with a
    54-frame halo (only the depthwise k7 convolutions mix frames: 3 taps x 3 blocks x 6 layers).  Same bits."""
    torch, AP, synth, sd, rest, _ = env
    from targetdiarization_b200 import Restorer
    x = synth.synthetic_fullband(2, 4 * 44100 + 17, seed=21).reshape(2, 1, -1).cuda()     # 401 frames per row
    full = rest(x)
    lib = rest._h.lib
    need = int(lib.tdz_apollo_workspace_bytes(2, x.shape[-1]))
    low = int(lib.tdz_apollo_min_workspace_bytes(2, x.shape[-1]))
    assert low < need // 1.5
    for budget in (low, (low + need) // 2):
        small = Restorer(sd, "cuda:0", max_workspace_bytes=budget)
        out = small(x)
        assert small._ws.numel() <= max(budget, low)
        assert torch.equal(out, full), budget


def test_restorer_from_pretrain_round_trip(env, tmp_path):
    """BaseModel.serialize layout (base_model.py:132-146) -> from_pretrain with the reference's keyword arguments
    (AudioProcessor.py:279); other architectures are refused."""
    torch, AP, synth, sd, rest, _ = env
    from targetdiarization_b200 import Restorer
    path = str(tmp_path / "pytorch_model.bin")
    torch.save({"model_name": "Apollo", "state_dict": sd, "model_args": {"n_sample_rate": 2},
                "infos": {"software_versions": {"torch_version": torch.__version__}}}, path)
    r2 = Restorer.from_pretrain(path, sr=44100, win=20, feature_dim=256, layer=6)
    r2.eval()
    r2.to("cuda:0")
    x = synth.synthetic_fullband(1, 4000, seed=2).reshape(1, 1, -1).cuda()
    assert torch.equal(r2(x), rest(x))
    with pytest.raises(ValueError):
        Restorer.from_pretrain(path, sr=44100, win=20, feature_dim=256, layer=4)
    torch.save({"model_name": "MossFormer2", "state_dict": sd}, path)
    with pytest.raises(ValueError):
        Restorer.from_pretrain(path, sr=44100, win=20, feature_dim=256, layer=6)
