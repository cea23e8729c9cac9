"""ctypes binding of libtdz.so (include/tdz.h).  The product path has no fallback: if the CUDA library is
missing or cannot be loaded this module raises, it never routes to a CPU implementation."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TDZ_LIB: development switch, another build of the same library (A/B timing of kernel variants on one box)
LIB_PATH = os.environ.get("TDZ_LIB") or os.path.join(_HERE, "libtdz.so")

NUM_LAYERS = 24
c_f32p = ctypes.c_void_p  # device pointers travel as integers


class LayerWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "w_in", "b_in", "dw_in", "os_gamma", "os_beta", "w_out", "b_out", "dw_out", "w_c1", "b_c1", "prelu_c1",
        "ln1_g", "ln1_b", "w_uv", "b_uv", "dw_uv", "w_lin", "b_lin", "w_proj", "dd_w1", "in1_g", "in1_b",
        "dd_prelu1", "dd_w2", "in2_g", "in2_b", "dd_prelu2", "w_c2", "b_c2")]


class MossFormer2Weights(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in (
        "enc_w", "w_enc1x1", "enc1x1_colsum", "enc1x1_bias", "pos_inv_freq", "pos_scale", "rot_freqs")]
        + [("layers", LayerWeights * NUM_LAYERS)]
        + [(n, ctypes.c_void_p) for n in (
            "fln_g", "fln_b", "fgn_g", "fgn_b", "mask_prelu", "w_out1", "b_out1", "w_tg", "b_tg", "w_dec1",
            "dec_w", "dec_wt")])


class SepLayout(ctypes.Structure):
    _names = ("enc", "x0", "x", "xbf", "ss", "vu", "qk4", "lq_lo", "qkf", "P", "o", "o_ss", "c", "nhat", "xuv", "xubf", "f1", "p", "y1",
              "y2", "g", "lnb", "ab", "mb", "gated", "sep", "kv_part", "kv", "gn_stats", "in_stats", "in_ss", "samp",
              "rot", "hrs", "pos", "total")
    _fields_ = ([(n, ctypes.c_size_t) for n in _names]
                + [("S", ctypes.c_int64), ("Sp", ctypes.c_int64), ("Mtot", ctypes.c_int64),
                   ("kv_nsplit", ctypes.c_int32), ("kv_kb_per_split", ctypes.c_int32)])


SV_NUM_BLOCKS = 16


class Conv(ctypes.Structure):
    _fields_ = [("w", ctypes.c_void_p), ("b", ctypes.c_void_p)]


class EresBlock(ctypes.Structure):
    _fields_ = [("conv1", Conv), ("convs", Conv * 4), ("aff_a", Conv * 3), ("aff_b", Conv * 3), ("conv3", Conv),
                ("shortcut", Conv)]


class Eres2NetV2Weights(ctypes.Structure):
    _fields_ = [("stem_w", ctypes.c_void_p), ("stem_b", ctypes.c_void_p), ("blocks", EresBlock * SV_NUM_BLOCKS),
                ("layer3_ds", Conv), ("fuse_a", Conv), ("fuse_b", Conv), ("seg1", Conv)]


# name -> (restype, argtypes); mirrors include/tdz.h one to one
_vp, _i64, _sz, _int, _f = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
AP_LAYERS = 6


class StftPlan(ctypes.Structure):
    _fields_ = [("n_fft", ctypes.c_int32), ("hop", ctypes.c_int32), ("window_dev", ctypes.c_void_p),
                ("twiddle_dev", ctypes.c_void_p)]


class ApolloIcb(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("dw", "dw_b", "w1", "b1", "w2", "b2")]


class ApolloLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w_qkv", "w_out", "w_mlp1", "w_mlp2")] + [("icb", ApolloIcb * 3)]


class ApolloWeights(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in ("bn_g", "bn_w", "bn_b", "rot_cos", "rot_sin")]
                + [("layers", ApolloLayer * AP_LAYERS)]
                + [(n, ctypes.c_void_p) for n in ("out_g", "out_wv", "out_wg", "out_bv", "out_bg")]
                + [("plan", StftPlan)])


SIGNATURES = {
    "tdz_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "tdz_destroy": (None, [_vp]),
    "tdz_last_error": (ctypes.c_char_p, [_vp]),
    "tdz_num_sms": (_int, [_vp]),
    "tdz_version": (ctypes.c_char_p, []),
    "tdz_set_mossformer2_weights": (_int, [_vp, ctypes.POINTER(MossFormer2Weights)]),
    "tdz_num_frames": (_i64, [_i64]),
    "tdz_padded_frames": (_i64, [_i64]),
    "tdz_separate_workspace_bytes": (_sz, [_i64, _i64]),
    "tdz_separate": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "tdz_separate_strided": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _sz, _vp]),
    "tdz_separate_debug": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _int, _int, _int]),
    "tdz_separate_layout": (_int, [_i64, _i64, _int, ctypes.POINTER(SepLayout)]),
    "tdz_gather_segments": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "tdz_gather_segments_span": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "tdz_stitch_ola": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f, _vp, _vp]),
    "tdz_stitch_concat": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "tdz_loudness_blocks": (_int, [_vp, _vp, _i64, _i64, ctypes.POINTER(ctypes.c_double), _vp, _vp, _i64,
                                   ctypes.c_double, _vp, _vp, _vp]),
    "tdz_fbank_frames": (_i64, [_i64]),
    "tdz_set_fbank_tables": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "tdz_fbank": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "tdz_set_eres2netv2_weights": (_int, [_vp, ctypes.POINTER(Eres2NetV2Weights)]),
    "tdz_embed_workspace_bytes": (_sz, [_i64, _i64]),
    "tdz_embed": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "tdz_embed_debug": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _int]),
    "tdz_cosine_scores": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "tdz_stft_frames": (_i64, [_i64, _i64]),
    "tdz_stft": (_int, [_vp, ctypes.POINTER(StftPlan), _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _vp]),
    "tdz_istft": (_int, [_vp, ctypes.POINTER(StftPlan), _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i64,
                         _vp]),
    "tdz_set_apollo_weights": (_int, [_vp, ctypes.POINTER(ApolloWeights)]),
    "tdz_apollo_workspace_bytes": (_sz, [_i64, _i64]),
    "tdz_apollo_min_workspace_bytes": (_sz, [_i64, _i64]),
    "tdz_apollo_restore": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "tdz_apollo_debug": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp, _int]),
}

_lib = None


def load():
    """Loads libtdz.so (built by __graft_entry__.build()).  Raises if it is absent -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "the tdz hot path has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def resolve_device(device):
    """torch.device with an explicit index: 'cuda' means the CURRENT device (like torch), not GPU 0."""
    import torch
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def free_device_bytes(device, held=0):
    """Bytes a workspace may grow to on `device`: what the driver reports free, plus what torch's caching allocator
    holds but does not use, plus the workspace the caller already owns (`held`, reused when it grows)."""
    import torch
    free, _ = torch.cuda.mem_get_info(device)
    cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return int(free + max(cached, 0) + held)


class CallGuard:
    """Serialises the calls into one host object (the reference shares ONE global model between
    the main thread and the ThreadPoolExecutor workers of its WebSocket server without any locking, main.py:42,338-367).

    A Separator / Embedder owns mutable state - the workspace its kernels scribble in, the static buffers of its
    captured CUDA graphs - so two Python threads must not interleave inside a call (lock), and a call on another CUDA
    stream than the previous one must not start before that one's kernels are done with the workspace (the new stream
    waits for the old one; nothing is recorded on the normal same-stream path)."""

    def __init__(self, device):
        import threading
        self.device = device
        self.lock = threading.RLock()
        self.depth = 0
        self.stream = None

    def __enter__(self):
        import torch
        self.lock.acquire()
        self.depth += 1
        if self.depth == 1:
            cur = torch.cuda.current_stream(self.device)
            last = self.stream
            if last is not None and last.cuda_stream != cur.cuda_stream and not torch.cuda.is_current_stream_capturing():
                cur.wait_stream(last)
            if last is None or last.cuda_stream != cur.cuda_stream:
                if not torch.cuda.is_current_stream_capturing():
                    self.stream = cur
        return self

    def __exit__(self, *exc):
        self.depth -= 1
        self.lock.release()
        return False


class Handle:
    """Owns one tdz_ctx.  check() turns non-zero return codes into RuntimeError (the reference's convention
    is Python exceptions caught at init, AudioProcessor.py:187-194)."""

    def __init__(self, device=0):
        self.lib = load()
        ptr = ctypes.c_void_p()
        rc = self.lib.tdz_create(int(device), ctypes.byref(ptr))
        if rc != 0 or not ptr.value:
            raise RuntimeError(f"tdz_create(device={device}) failed with code {rc}: an sm_100 GPU is required "
                               "(no CPU fallback)")
        self.ptr = ptr
        self.device = int(device)

    def check(self, rc, what):
        if rc != 0:
            msg = self.lib.tdz_last_error(self.ptr)
            raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc}")

    @property
    def num_sms(self):
        return self.lib.tdz_num_sms(self.ptr)

    def close(self):
        if getattr(self, "ptr", None) is not None and self.ptr.value:
            self.lib.tdz_destroy(self.ptr)
            self.ptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
