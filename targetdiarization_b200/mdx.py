"""`ConvTDFNet`: the STFT / inverse-STFT front and back end the reference wraps around its MDX-Net ONNX model
(AudioProcessor.py:65-120, built at :241, used by `denoise_vocal` at :624,:632).  Same constructor and methods:
`stft(x[B, 2, chunk]) -> [B, 4, dim_f, dim_t]`, `istft(x[B, 4, dim_f, dim_t]) -> [B, 2, chunk]` (returned on the host
like the reference's `.cpu()`).  The transforms run in libtdz.so (tdz_stft / tdz_istft: shared-memory mixed-radix FFT,
two real frames per complex transform); there is no CPU path."""
import torch

from . import _lib
from .weights import make_stft_plan


class ConvTDFNet:
    def __init__(self, target_name, L, dim_f, dim_t, n_fft, hop=1024, device="cuda:0", handle=None):
        self.dim_c = 4
        self.dim_f = dim_f
        self.dim_t = 2 ** dim_t
        self.n_fft = n_fft
        self.hop = hop
        self.device = torch.device(device if device else "cuda:0")
        if self.device.type != "cuda":
            raise RuntimeError("tdz.ConvTDFNet runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self.device = _lib.resolve_device(self.device)
        self.n_bins = self.n_fft // 2 + 1
        self.chunk_size = hop * (self.dim_t - 1)
        self.target_name = target_name
        self.n = L // 2
        if not 0 < dim_f <= self.n_bins:
            raise ValueError(f"dim_f {dim_f} outside 1 .. n_fft / 2 + 1 = {self.n_bins}")
        self._h = handle if handle is not None else _lib.Handle(self.device.index)
        self._plan, self._keep = make_stft_plan(n_fft, hop, self.device)
        self.window = self._keep[0]
        self._frames = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def stft(self, x):
        x = torch.as_tensor(x).to(self.device, torch.float32).reshape(-1, self.chunk_size).contiguous()
        rows = x.shape[0]
        if rows % 2:
            raise ValueError("ConvTDFNet.stft takes stereo chunks [B, 2, chunk_size]")
        F, T = self.dim_f, self.dim_t
        out = torch.empty(rows // 2, self.dim_c, F, T, dtype=torch.float32, device=self.device)
        import ctypes
        self._h.check(self._h.lib.tdz_stft(self._h.ptr, ctypes.byref(self._plan), x.data_ptr(), rows, self.chunk_size, F,
                                           out.data_ptr(), 2 * F * T, T, 1, F * T, self._stream()), "tdz_stft")
        return out

    def istft(self, x, freq_pad=None):
        """freq_pad: the reference concatenates zeros for the bins above dim_f (or a caller-supplied tensor); zeros are
        implicit here, a non-zero freq_pad is concatenated on the device first."""
        import ctypes
        x = torch.as_tensor(x).to(self.device, torch.float32)
        F = self.dim_f
        if freq_pad is not None:
            x = torch.cat([x, torch.as_tensor(freq_pad).to(self.device, torch.float32)], -2)
            F = x.shape[-2]
        x = x.contiguous()
        c = 4 * 2 if self.target_name == "*" else 2
        T = x.shape[-1]
        rows = x.numel() // (2 * F * T)
        need = rows * T * self.n_fft
        if self._frames is None or self._frames.numel() < need:
            self._frames = torch.empty(need, dtype=torch.float32, device=self.device)
        out_len = self.hop * (T - 1)
        out = torch.empty(rows, out_len, dtype=torch.float32, device=self.device)
        self._h.check(self._h.lib.tdz_istft(self._h.ptr, ctypes.byref(self._plan), x.data_ptr(), rows, T, F,
                                            2 * F * T, T, 1, F * T, self._frames.data_ptr(), out.data_ptr(), out_len,
                                            self._stream()), "tdz_istft")
        return out.reshape(-1, c, self.chunk_size).cpu()
