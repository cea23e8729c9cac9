"""Test stand-in for targetdiarization_b200.pipeline.CudaKernels on CPU tensors, used ONLY by the gloo tests of
the multi-rank sharding / gather logic (no GPU in the build container).  It restates the semantics of the
tdz_gather_segments_span / tdz_stitch_ola / tdz_separate_strided entry points with torch CPU ops and takes the
separator as a callable.  It is test infrastructure, not a product path."""
import numpy as np
import torch


class NumpyKernels:
    def __init__(self, model, max_batch=4):
        self.model = model
        self._max_batch = max_batch
        self.device = torch.device("cpu")
        self.calls = []
        self.uploaded = []   # lengths of the host arrays handed to to_device (how much each rank uploads)

    def to_device(self, audio):
        if isinstance(audio, np.ndarray):
            self.uploaded.append(int(audio.shape[0]))
        return torch.as_tensor(audio, dtype=torch.float32).contiguous()

    def to_host(self, dev):
        return dev.numpy()

    def empty(self, *shape):
        return torch.full(shape, float("nan"), dtype=torch.float32)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32)

    def max_batch(self, T):
        return self._max_batch

    def separate(self, chunks, out=None, out_strides=None):
        n, T = chunks.shape
        self.calls.append((n, T))
        est = torch.cat([self.model(chunks[i:i + 1]) for i in range(n)], 0)   # [n,2,T]
        if out is None:
            return est
        cs, ss = out_strides if out_strides is not None else (2 * T, T)
        flat = out.reshape(-1)
        assert flat.data_ptr() == out.data_ptr(), "out must be a view (written in place)"
        for k in range(n):
            for s in range(2):
                flat[k * cs + s * ss:k * cs + s * ss + T] = est[k, s]
        return out

    def gather_segments(self, mix, plan, seg_lo, n_seg, mix_origin=0):
        seg = torch.zeros(n_seg, plan.session)
        for j in range(n_seg):
            a, b = plan.segment_range(seg_lo + j)
            lo, hi = max(a, 0), min(b, plan.length)
            if hi > lo:
                assert lo >= mix_origin and hi <= mix_origin + mix.shape[0], "segment reads non-resident samples"
                seg[j, lo - a:hi - a] = mix[lo - mix_origin:hi - mix_origin]
        return seg

    def stitch_ola(self, est, plan, seg_lo, out_begin, n_out, out=None):
        acc = torch.zeros(2, n_out)
        for j in range(est.shape[0]):  # ascending segment order, like the kernel
            a, _ = plan.segment_range(seg_lo + j)
            lo, hi = max(a, out_begin), min(a + plan.session, out_begin + n_out, plan.length)
            if hi > lo:
                acc[:, lo - out_begin:hi - out_begin] += est[j, :, lo - a:hi - a]
        acc = acc / plan.ratio
        if out is None:
            return acc
        out.copy_(acc)
        return out
