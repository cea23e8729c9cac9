"""`Restorer`: the object the reference keeps in `AudioProcessor.restorer` (AudioProcessor.py:276-281:
`BaseModel.from_pretrain(weights_file, sr=44100, win=20, feature_dim=256, layer=6)`) and calls as
`self.restorer(tensor[1, nch, nsample])` in `restore_audio` (AudioProcessor.py:959-980).  Same call surface as the
reference `Apollo` module (look2hear/models/apollo.py:278-297): float32 [B, nch, nsample] in, [B, nch, nsample] out on
the same CUDA device.  The compute is libtdz.so (tdz_apollo_restore: shared-memory FFT, tcgen05 GEMMs for the 1x1
convolutions, SIMT kernels for the band split / merge, attention and depthwise convolutions); there is no CPU path."""
import ctypes

import numpy as np
import torch

from . import _lib
from .weights import APOLLO_ARGS, PackedApollo, check_apollo_state_dict

HOP = 441
TAPS = {"spec": 0, "feat": 1, "att0": 2, "band0": 3, "layer0": 4, "layer1": 5, "layer2": 6, "layer3": 7, "layer4": 8,
        "layer5": 9, "est_spec": 10}


def check_apollo_args(*args, **kwargs):
    """Apollo.__init__(sr, win, feature_dim, layer) (apollo.py:215-221): the kernels implement the one configuration
    the reference instantiates (AudioProcessor.py:279); anything else is refused, never ignored."""
    names = [n for n, _ in APOLLO_ARGS]
    if len(args) > len(names):
        raise TypeError(f"Apollo takes at most {len(names)} positional arguments, got {len(args)}")
    given = dict(zip(names, args))
    for k, v in kwargs.items():
        if k not in names:
            raise TypeError(f"Apollo.__init__() got an unexpected keyword argument {k!r}")
        if k in given:
            raise TypeError(f"Apollo.__init__() got multiple values for argument {k!r}")
        given[k] = v
    for k, want in APOLLO_ARGS:
        if k in given and given[k] != want:
            raise ValueError(f"tdz.Restorer implements Apollo({k}={want!r}) only; asked for {k}={given[k]!r}")


class Restorer:
    sample_rate = 44100
    # launches of one forward (csrc/ap_api.cuh): stft, band split; per layer qkv GEMM + attention + 3 GEMMs +
    # 3 x (dwconv + back-to-back GEMM) = 11; band merge; istft (2)
    KERNELS_PER_FORWARD = 2 + 6 * 11 + 1 + 2

    def __init__(self, state_dict=None, device="cuda:0", handle=None, max_workspace_bytes=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tdz.Restorer runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self.device = _lib.resolve_device(self.device)
        self._h = handle if handle is not None else _lib.Handle(self.device.index)
        self._packed = None
        self._ws = None
        self._ws_raw = None
        # None: up to 60 % of the memory that is free when a call is sized.  A single pass needs 493 KB per 10 ms frame
        # (an hour of audio: 180 GB); with less the library runs the network over frame chunks with a 54-frame halo -
        # the receptive field of the depthwise convolutions - which gives the same bits.
        self.max_workspace_bytes = None if max_workspace_bytes is None else int(max_workspace_bytes)
        self._guard = _lib.CallGuard(self.device)    # one call at a time per object, any Python thread / stream
        if state_dict is not None:
            self.load_state_dict(state_dict)

    @classmethod
    def from_pretrain(cls, pretrained_model_conf_or_path, *args, device="cuda:0", **kwargs):
        """BaseModel.from_pretrain (base_model.py:118-130) on a file in BaseModel.serialize layout (:132-146)."""
        try:
            conf = torch.load(pretrained_model_conf_or_path, map_location="cpu", weights_only=True)
        except Exception:
            conf = torch.load(pretrained_model_conf_or_path, map_location="cpu", weights_only=False)
        if "model_name" not in conf or "state_dict" not in conf:
            raise KeyError("checkpoint lacks 'model_name' / 'state_dict' (BaseModel.serialize layout)")
        if conf["model_name"] != "Apollo":
            raise ValueError(f"tdz.Restorer implements Apollo only, checkpoint holds {conf['model_name']}")
        check_apollo_args(*args, **kwargs)
        return cls(conf["state_dict"], device)

    def load_state_dict(self, state_dict, strict=True):
        check_apollo_state_dict(state_dict, strict=strict)
        self._packed = PackedApollo(state_dict, self.device)
        self._h.check(self._h.lib.tdz_set_apollo_weights(self._h.ptr, ctypes.byref(self._packed.table)),
                      "tdz_set_apollo_weights")
        return self

    def eval(self):
        return self

    def to(self, device):
        if _lib.resolve_device(device) != self.device:
            raise RuntimeError("tdz.Restorer is bound to the device it was created on")
        return self

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = self._ws_raw = None
            self._ws_raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-self._ws_raw.data_ptr()) % 1024
            self._ws = self._ws_raw[off:off + nbytes]
        return self._ws

    def _run(self, x, tap=None):
        if self._packed is None:
            raise RuntimeError("Restorer has no weights; call load_state_dict first")
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.ndim != 3:
            raise ValueError(f"Apollo.forward takes [B, nch, nsample], got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        B, nch, ns = x.shape
        rows = B * nch
        if ns <= HOP:
            raise ValueError(f"input of {ns} samples is too short for the reflect-padded STFT (needs more than {HOP})")
        lib, h = self._h.lib, self._h
        nbytes = int(lib.tdz_apollo_workspace_bytes(rows, ns))
        held = self._ws.numel() if self._ws is not None else 0
        if tap is None and nbytes > held:      # the workspace would have to grow: is there room for a single pass?
            budget = self.max_workspace_bytes
            if budget is None:                 # (a driver query: only when a call does not fit what is already held)
                budget = int(0.6 * _lib.free_device_bytes(self.device, held))
            if nbytes > budget:
                nbytes = max(int(lib.tdz_apollo_min_workspace_bytes(rows, ns)), min(budget, nbytes), held)
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        T = 1 + ns // HOP
        if tap is None:
            out = torch.empty(B, nch, ns, dtype=torch.float32, device=self.device)
            h.check(lib.tdz_apollo_restore(h.ptr, x.data_ptr(), rows, ns, out.data_ptr(), ws.data_ptr(), nbytes, stream),
                    "tdz_apollo_restore")
            return out
        shape = (rows, T, 442, 2) if tap in ("spec", "est_spec") else (rows, T, 80, 256)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        h.check(lib.tdz_apollo_debug(h.ptr, x.data_ptr(), rows, ns, out.data_ptr(), ws.data_ptr(), nbytes, stream,
                                     TAPS[tap]), "tdz_apollo_debug")
        return out

    def __call__(self, x):
        with self._guard:
            return self._run(x)

    forward = __call__

    def tap(self, x, name):
        """Test hook: an intermediate of the forward in the oracle's layout (oracle/apollo_port.py `taps`)."""
        with self._guard:
            return self._run(x, name)
