"""CPU tests of the drop-in boundary: libtdz.so loads without a GPU, exports every symbol include/tdz.h declares,
the ctypes table mirrors the header, and the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tdz.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\*?\s*(tdz_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from targetdiarization_b200 import _lib
    return _lib


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("tdz_create", "tdz_separate", "tdz_stitch_ola", "tdz_stitch_concat", "tdz_gather_segments", "tdz_separate_strided",
                 "tdz_gather_segments_span",
                 "tdz_fbank", "tdz_embed", "tdz_cosine_scores", "tdz_last_error", "tdz_stft", "tdz_istft",
                 "tdz_set_apollo_weights", "tdz_apollo_restore"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol(lib):
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(so, name), f"{name} declared in include/tdz.h but not exported by libtdz.so"


def test_ctypes_table_matches_header(lib):
    assert sorted(lib.SIGNATURES) == declared_functions()


def test_layout_helpers_without_gpu(lib):
    so = lib.load()
    assert so.tdz_version().startswith(b"tdz")
    assert so.tdz_num_frames(64000) == 7999 and so.tdz_padded_frames(64000) == 8192
    assert so.tdz_num_frames(138634) == 17328 and so.tdz_num_frames(15) == 0
    assert so.tdz_fbank_frames(64000) == 398 and so.tdz_fbank_frames(399) == 0
    lay = lib.SepLayout()
    assert so.tdz_separate_layout(64, 64000, 148, ctypes.byref(lay)) == 0
    assert (lay.S, lay.Sp, lay.Mtot) == (7999, 8192, 64 * 8192)
    assert lay.total % 1024 == 0 and lay.total > lay.Mtot * 16384


def test_restorer_tables_without_gpu(lib):
    so = lib.load()
    assert so.tdz_stft_frames(261120, 1024) == 256 and so.tdz_stft_frames(22173, 441) == 51
    assert so.tdz_apollo_workspace_bytes(1, 441) == 0          # too short for the reflect padding
    n = so.tdz_apollo_workspace_bytes(2, 44100)                # 2 rows x 101 frames x 80 bands = 16 160 tokens
    assert n % 1024 == 0 and n > 16160 * 6000
    assert so.tdz_apollo_min_workspace_bytes(2, 44100) == n    # shorter than the smallest chunk: one pass
    hour = 3600 * 44100
    assert so.tdz_apollo_workspace_bytes(1, hour) > 170e9 and so.tdz_apollo_min_workspace_bytes(1, hour) < 4.5e9
    # tdz_stft_plan = 2 int32 + 2 pointers; per layer 4 + 3*6 pointers; model 5 + 6*22 + 5 pointers + the plan
    assert ctypes.sizeof(lib.StftPlan) == 24
    assert ctypes.sizeof(lib.ApolloLayer) == 22 * 8
    assert ctypes.sizeof(lib.ApolloWeights) == (5 + 6 * 22 + 5) * 8 + 24


def test_struct_sizes_match_header(lib):
    # pointer tables only: 29 pointers per layer; 7 + 24*29 + 12 for the model
    assert ctypes.sizeof(lib.LayerWeights) == 29 * 8
    assert ctypes.sizeof(lib.MossFormer2Weights) == (7 + 24 * 29 + 12) * 8
    assert ctypes.sizeof(lib.EresBlock) == (1 + 4 + 3 + 3 + 1 + 1) * 16
    assert ctypes.sizeof(lib.Eres2NetV2Weights) == 16 + 16 * ctypes.sizeof(lib.EresBlock) + 4 * 16


def test_no_cpu_fallback(lib):
    import torch
    from targetdiarization_b200 import Embedder, Separator
    with pytest.raises(RuntimeError):
        Separator(None, "cpu")
    with pytest.raises(RuntimeError):
        Embedder(None, "cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            lib.Handle(0)   # tdz_create refuses anything but an sm_100 GPU


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "targetdiarization_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_from_pretrain_argument_validation():
    """`from_pretrain(path, **cfg.model)` (AudioProcessor.py:269-273): the constructor arguments of the reference
    MossFormer2 (mossformer2.py:532-541) are validated, never silently ignored."""
    from targetdiarization_b200.separator import MODEL_ARGS, check_model_args
    check_model_args()
    check_model_args(**dict(MODEL_ARGS))
    check_model_args(512, 512, 24, 16, "ln", 2)
    for bad in (dict(num_blocks=12), dict(in_channels=256), dict(num_spks=3), dict(norm="gln"),
                dict(skip_around_intra=False), dict(kernel_size=8)):
        with pytest.raises(ValueError):
            check_model_args(**bad)
    with pytest.raises(TypeError):
        check_model_args(sample_rate=16000)          # MossFormer2.__init__ has no such argument either
    with pytest.raises(TypeError):
        check_model_args(512, in_channels=512)
    with pytest.raises(ValueError):
        check_model_args(512, 512, 6)


def test_state_dict_contract():
    """load_state_dict(strict=True) semantics for the reference's 1 099-key dict."""
    from targetdiarization_b200 import synth, weights
    sd = synth.random_state_dict(seed=0)
    assert len(weights.mossformer2_key_shapes()) == 1099
    weights.check_mossformer2_state_dict(sd)
    extra = dict(sd, bogus=sd["dec.weight"])
    with pytest.raises(RuntimeError, match="Unexpected"):
        weights.check_mossformer2_state_dict(extra)
    weights.check_mossformer2_state_dict(extra, strict=False)
    with pytest.raises(RuntimeError, match="Missing"):
        weights.check_mossformer2_state_dict({k: v for k, v in sd.items() if k != "dec.weight"})
    with pytest.raises(RuntimeError, match="size mismatch"):
        weights.check_mossformer2_state_dict(dict(sd, **{"dec.weight": sd["dec.weight"][:256]}))


def test_embedder_oracle_owns_its_architecture_table():
    """The ERes2NetV2 oracle's block table is its own (not imported from the product) and the shared weight
    generator agrees with it key for key."""
    import inspect
    from oracle import eres2netv2_port as E
    from targetdiarization_b200 import synth
    assert "import block_specs" not in inspect.getsource(E)
    assert [tuple(r) for r in E.block_specs()] == [tuple(r) for r in synth.block_specs()]
    sd = E.random_state_dict(seed=1)      # raises if the generator's keys / shapes differ from the oracle's table
    assert len(sd) == len(E.expected_key_shapes())
