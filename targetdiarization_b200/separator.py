"""`Separator`: the object the reference keeps in `AudioProcessor.separater` (AudioProcessor.py:268-274) and
calls as `self.separater(tensor)` (:943).  Same call surface as the reference `MossFormer2` module
(look2hear/models/mossformer2.py:563-589): accepts [T], [B,T] or [B,1,T] float32, returns [B,2,T] float32 on
the same CUDA device.  The compute is libtdz.so (hand-written sm_100a kernels); there is no CPU path."""
import ctypes

import torch

from . import _lib
from .weights import PackedMossFormer2


# MossFormer2.__init__ (look2hear/models/mossformer2.py:532-541), in positional order.  The kernels are written for
# exactly this architecture (tile shapes, channel counts and the 24-layer weight table of include/tdz.h are
# compile-time); any other value is refused instead of silently ignored.
MODEL_ARGS = (("in_channels", 512), ("out_channels", 512), ("num_blocks", 24), ("kernel_size", 16), ("norm", "ln"),
              ("num_spks", 2), ("skip_around_intra", True), ("use_global_pos_enc", True), ("max_length", 20000))


def check_model_args(*args, **kwargs):
    """Validates the constructor arguments the reference passes through `from_pretrain(path, **cfg.model)`
    (AudioProcessor.py:269-273, base_model.py:118-130) against the one architecture libtdz.so implements."""
    names = [n for n, _ in MODEL_ARGS]
    if len(args) > len(names):
        raise TypeError(f"MossFormer2 takes at most {len(names)} positional arguments, got {len(args)}")
    given = dict(zip(names, args))
    for k, v in kwargs.items():
        if k not in names:
            raise TypeError(f"MossFormer2.__init__() got an unexpected keyword argument {k!r}")
        if k in given:
            raise TypeError(f"MossFormer2.__init__() got multiple values for argument {k!r}")
        given[k] = v
    for k, want in MODEL_ARGS:
        if k in given and given[k] != want:
            raise ValueError(f"tdz.Separator implements MossFormer2({k}={want!r}) only; the checkpoint / config asks "
                             f"for {k}={given[k]!r}")


# Calls of at most this many padded frames (B * Sp) are launch-bound: ~470 kernels of a few microseconds each.  From the
# second call of a shape on they replay a CUDA graph captured from the same launch sequence (one submission instead of
# ~470; same kernels, same bits).  32 768 frames = 16 concurrent 10 s windows or 25 concurrent 600 ms chunks.
GRAPH_MAX_FRAMES = 32768
GRAPH_CACHE = 16


class Separator:
    sample_rate = 16000
    num_spks = 2

    def __init__(self, state_dict=None, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tdz.Separator runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self.device = _lib.resolve_device(self.device)
        self._h = _lib.Handle(self.device.index)
        self._packed = None
        self._ws = None
        self._ws_generation = 0      # bumped when the workspace moves: captured CUDA graphs hold its address
        self._graphs = {}            # (B, T) -> [times seen, CUDAGraph | None, static input, static output]
        self.graph_max_frames = GRAPH_MAX_FRAMES
        self._guard = _lib.CallGuard(self.device)    # one call at a time per object, any Python thread / stream
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -- nn.Module-like surface the reference touches (base_model.py:118-130; AudioProcessor.py:271-273)
    @classmethod
    def from_state_dict(cls, state_dict, device="cuda:0"):
        return cls(state_dict, device)

    @classmethod
    def from_pretrain(cls, pretrained_model_conf_or_path, *args, device="cuda:0", **kwargs):
        """BaseModel.from_pretrain (base_model.py:118-130): the file holds what BaseModel.serialize wrote (:132-146) -
        {'model_name', 'state_dict', 'model_args', 'infos'}; *args / **kwargs are the constructor arguments the
        reference forwards from its config.yaml (`from_pretrain(path, **cfg.model)`, AudioProcessor.py:269-273).
        Anything but the default MossFormer2 architecture is refused (ValueError / TypeError), never ignored."""
        try:
            conf = torch.load(pretrained_model_conf_or_path, map_location="cpu", weights_only=True)
        except Exception:
            # BaseModel.serialize stores version objects under 'infos' (base_model.py:141-145), which the tensors-only
            # loader refuses; the reference itself unpickles the whole file (torch.load without restrictions)
            conf = torch.load(pretrained_model_conf_or_path, map_location="cpu", weights_only=False)
        if "model_name" not in conf or "state_dict" not in conf:
            raise KeyError("checkpoint lacks 'model_name' / 'state_dict' (BaseModel.serialize layout)")
        if conf["model_name"] != "MossFormer2":
            raise ValueError(f"tdz implements MossFormer2 only, checkpoint holds {conf['model_name']}")
        check_model_args(*args, **kwargs)
        if isinstance(conf.get("model_args"), dict):
            check_model_args(**conf["model_args"])
        return cls(conf["state_dict"], device)

    def load_state_dict(self, state_dict, strict=True):
        """Same contract as nn.Module.load_state_dict(strict=True) for the reference's 1 099-key dict: missing /
        unexpected keys and shape mismatches raise RuntimeError (a checkpoint of another architecture must not load)."""
        from .weights import check_mossformer2_state_dict
        check_mossformer2_state_dict(state_dict, strict=strict)
        self._packed = PackedMossFormer2(state_dict, self.device)
        self._h.check(self._h.lib.tdz_set_mossformer2_weights(self._h.ptr, ctypes.byref(self._packed.table)),
                      "tdz_set_mossformer2_weights")
        return self

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("tdz.Separator is bound to the device it was created on")
        return self

    # -- forward
    def workspace_bytes(self, B, T):
        return int(self._h.lib.tdz_separate_workspace_bytes(B, T))

    def _workspace(self, nbytes):
        """Caller-owned scratch, grown on demand; 1024 B aligned (TMA / swizzle atoms), which the caching allocator
        alone does not guarantee (512 B)."""
        if self._ws is None or self._ws.numel() < nbytes:
            self._graphs.clear()
            self._ws_generation += 1
            self._ws = None
            self._ws_raw = None
            self._ws_raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self._ws_raw.data_ptr()) % 1024
            self._ws = self._ws_raw[off:off + nbytes]
        return self._ws

    def _graphed(self, x, B, T):
        """Small calls: eager the first time a shape is seen, captured into a CUDA graph the second time, replayed
        from then on (the library only launches on the caller's stream and never allocates, so its launch sequence
        is capturable as it is).  Returns None when this call should take the eager path."""
        key = (B, T)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= GRAPH_CACHE:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = [1, None, None, None]
            return None
        if ent[1] is None:
            nbytes = self.workspace_bytes(B, T)
            ws = self._workspace(nbytes)
            ent = self._graphs.setdefault(key, [1, None, None, None])   # _workspace may have cleared the cache
            s_in = torch.empty(B, T, dtype=torch.float32, device=self.device)
            s_out = torch.empty(B, 2, T, dtype=torch.float32, device=self.device)
            lib, h = self._h.lib, self._h
            g = torch.cuda.CUDAGraph()
            torch.cuda.current_stream(self.device).synchronize()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):  # other threads may allocate meanwhile
                stream = torch.cuda.current_stream(self.device).cuda_stream
                rc = lib.tdz_separate(h.ptr, s_in.data_ptr(), B, T, s_out.data_ptr(), ws.data_ptr(), nbytes, stream)
            h.check(rc, "tdz_separate (graph capture)")
            ent[1], ent[2], ent[3] = g, s_in, s_out
        ent[0] += 1
        ent[2].copy_(x)
        ent[1].replay()
        return ent[3].clone()


    def layout(self, B, T):
        lay = _lib.SepLayout()
        rc = self._h.lib.tdz_separate_layout(B, T, self._h.num_sms, ctypes.byref(lay))
        if rc:
            raise RuntimeError("tdz_separate_layout failed")
        return lay

    def __call__(self, mix, _debug=None, out=None, out_strides=None):
        """mix [T] / [B,T] / [B,1,T] -> [B,2,T].  `out` (optional) = a caller-owned fp32 device buffer to write
        into; with `out_strides=(chunk_stride, speaker_stride)` (in floats) stream s of chunk b goes to
        out.data_ptr() + 4*(b*chunk_stride + s*speaker_stride) - how the chunk loop writes windows straight into
        the stitched [2, L] streams."""
        if self._packed is None:
            raise RuntimeError("Separator has no weights; call load_state_dict first")
        with self._guard:
            return self._forward(mix, _debug, out, out_strides)

    def _forward(self, mix, _debug, out, out_strides):
        x = mix
        if x.ndim == 1:
            x = x.unsqueeze(0)
        if x.ndim == 3:
            x = x.squeeze(1)
        if x.device != self.device:
            raise RuntimeError(f"input is on {x.device}, separator on {self.device}")
        x = x.to(torch.float32).contiguous()
        B, T = x.shape
        if (out is None and _debug is None and self.graph_max_frames
                and B * int(self._h.lib.tdz_padded_frames(T)) <= self.graph_max_frames and T >= 16
                and not torch.cuda.is_current_stream_capturing()):
            y = self._graphed(x, B, T)
            if y is not None:
                return y
        if out is None:
            out = torch.empty(B, 2, T, dtype=torch.float32, device=self.device)
        elif out.dtype != torch.float32 or out.device != self.device:
            raise RuntimeError("`out` must be a float32 tensor on the separator's device")
        nbytes = self.workspace_bytes(B, T)
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        lib, h = self._h.lib, self._h
        if _debug is None and out_strides is not None:
            rc = lib.tdz_separate_strided(h.ptr, x.data_ptr(), B, T, out.data_ptr(), int(out_strides[0]),
                                          int(out_strides[1]), ws.data_ptr(), nbytes, stream)
        elif _debug is None:
            if out.numel() < B * 2 * T or not out.is_contiguous():
                raise RuntimeError("`out` must be contiguous and hold [B,2,T]")
            rc = lib.tdz_separate(h.ptr, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), nbytes, stream)
        else:
            rc = lib.tdz_separate_debug(h.ptr, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), nbytes, stream,
                                        int(_debug[0]), int(_debug[1]), int(_debug[2]))
        h.check(rc, "tdz_separate")
        return out

    forward = __call__

    STEP_NAMES = ("ENCODER", "ENC1X1", "FLASH_IN", "SIM", "KV", "ATT_OUT", "TO_OUT", "FSMN_C1", "FSMN_UV", "FSMN_LIN",
                  "FSMN_PROJ", "DD1", "DD2", "FSMN_TAIL", "FSMN_C2", "FINAL_LN", "FINAL_GN", "OUT1", "TANHSIG", "DEC1",
                  "DECODER")
    LAYER_STEPS = STEP_NAMES[2:15]
    KERNELS_PER_FORWARD = 5 + 24 * 18 + 9  # launches of tdz_separate (csrc/tdz_api.cu), memsets not counted

    def time_steps(self, mix, reps=5):
        """CUDA-event time (ms) of every launch step of the forward run alone (layer 0 instance), after a
        full forward has populated the workspace.  Used by bench.py for the roofline of the dominant kernel."""
        self(mix)
        torch.cuda.synchronize(self.device)
        out = {}
        for k, name in enumerate(self.STEP_NAMES):
            # fsmn.linear -> project run as ONE back-to-back GEMM in the forward: "FSMN_LIN" is that fused launch
            # (steps k..k+1); "FSMN_PROJ" alone is the project kernel of the two-kernel form (tests only)
            hi = k + 1 if name == "FSMN_LIN" else k
            self(mix, _debug=(1, k, hi))  # warm
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                self(mix, _debug=(1, k, hi))
            e1.record()
            torch.cuda.synchronize(self.device)
            out[name] = e0.elapsed_time(e1) / reps
        return out

    def debug_buffer(self, B, T, name, dtype, cols):
        """Test hook: view of a named intermediate inside the workspace after a call ([B, Sp, cols])."""
        lay = self.layout(B, T)
        off = getattr(lay, name)
        es = torch.empty(0, dtype=dtype).element_size()
        n = lay.Mtot * cols
        return self._ws[off:off + n * es].view(dtype).view(B, lay.Sp, cols)
