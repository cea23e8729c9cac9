"""`Separator`: the object the reference keeps in `AudioProcessor.separater` (AudioProcessor.py:268-274) and
calls as `self.separater(tensor)` (:943).  Same call surface as the reference `MossFormer2` module
(look2hear/models/mossformer2.py:563-589): accepts [T], [B,T] or [B,1,T] float32, returns [B,2,T] float32 on
the same CUDA device.  The compute is libtdz.so (hand-written sm_100a kernels); there is no CPU path."""
import ctypes

import torch

from . import _lib
from .weights import PackedMossFormer2


class Separator:
    sample_rate = 16000
    num_spks = 2

    def __init__(self, state_dict=None, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tdz.Separator runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self._h = _lib.Handle(self.device.index or 0)
        self._packed = None
        self._ws = None
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # -- nn.Module-like surface the reference touches (base_model.py:118-130; AudioProcessor.py:271-273)
    @classmethod
    def from_state_dict(cls, state_dict, device="cuda:0"):
        return cls(state_dict, device)

    @classmethod
    def from_pretrain(cls, pretrained_model_conf_or_path, device="cuda:0", **kwargs):
        """Checkpoint layout of BaseModel.serialize (base_model.py:132-146): {'model_name','state_dict',...}."""
        conf = torch.load(pretrained_model_conf_or_path, map_location="cpu")
        if conf.get("model_name", "MossFormer2") != "MossFormer2":
            raise ValueError(f"tdz implements MossFormer2 only, checkpoint holds {conf['model_name']}")
        return cls(conf["state_dict"], device)

    def load_state_dict(self, state_dict, strict=True):
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
            self._ws = None
            self._ws_raw = None
            self._ws_raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self._ws_raw.data_ptr()) % 1024
            self._ws = self._ws_raw[off:off + nbytes]
        return self._ws

    def layout(self, B, T):
        lay = _lib.SepLayout()
        rc = self._h.lib.tdz_separate_layout(B, T, self._h.num_sms, ctypes.byref(lay))
        if rc:
            raise RuntimeError("tdz_separate_layout failed")
        return lay

    def __call__(self, mix, _debug=None):
        if self._packed is None:
            raise RuntimeError("Separator has no weights; call load_state_dict first")
        x = mix
        if x.ndim == 1:
            x = x.unsqueeze(0)
        if x.ndim == 3:
            x = x.squeeze(1)
        if x.device != self.device:
            raise RuntimeError(f"input is on {x.device}, separator on {self.device}")
        x = x.to(torch.float32).contiguous()
        B, T = x.shape
        out = torch.empty(B, 2, T, dtype=torch.float32, device=self.device)
        nbytes = self.workspace_bytes(B, T)
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        lib, h = self._h.lib, self._h
        if _debug is None:
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
    KERNELS_PER_FORWARD = 4 + 24 * 19 + 9  # launches of tdz_separate (csrc/tdz_api.cu), memsets not counted

    def time_steps(self, mix, reps=5):
        """CUDA-event time (ms) of every launch step of the forward run alone (layer 0 instance), after a
        full forward has populated the workspace.  Used by bench.py for the roofline of the dominant kernel."""
        self(mix)
        torch.cuda.synchronize(self.device)
        out = {}
        for k, name in enumerate(self.STEP_NAMES):
            self(mix, _debug=(1, k, k))  # warm
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                self(mix, _debug=(1, k, k))
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
