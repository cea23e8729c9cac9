"""Test stand-in for targetdiarization_b200.pipeline.CudaKernels on CPU tensors, used ONLY by the gloo tests of
the multi-rank sharding / gather logic (no GPU in the build container).  It restates the semantics of the
tdz_gather_segments / tdz_stitch_ola / tdz_stitch_concat kernels with torch CPU ops and takes the separator as
a callable.  It is test infrastructure, not a product path."""
import torch


class NumpyKernels:
    def __init__(self, model, max_batch=4):
        self.model = model
        self._max_batch = max_batch
        self.device = torch.device("cpu")
        self.calls = []

    def to_device(self, audio):
        return torch.as_tensor(audio, dtype=torch.float32).contiguous()

    def empty(self, *shape):
        return torch.full(shape, float("nan"), dtype=torch.float32)

    def max_batch(self, T):
        return self._max_batch

    def separate(self, chunks):
        self.calls.append(tuple(chunks.shape))
        return torch.cat([self.model(chunks[i:i + 1]) for i in range(chunks.shape[0])], 0)

    def gather_segments(self, mix, plan, seg_lo, n_seg):
        seg = torch.zeros(n_seg, plan.session)
        for j in range(n_seg):
            a, b = plan.segment_range(seg_lo + j)
            lo, hi = max(a, 0), min(b, plan.length)
            if hi > lo:
                seg[j, lo - a:hi - a] = mix[lo:hi]
        return seg

    def stitch_ola(self, est, plan, seg_lo, out_begin, n_out):
        out = torch.zeros(2, n_out)
        for j in range(est.shape[0]):  # ascending segment order, like the kernel
            a, _ = plan.segment_range(seg_lo + j)
            lo, hi = max(a, out_begin), min(a + plan.session, out_begin + n_out, plan.length)
            if hi > lo:
                out[:, lo - out_begin:hi - out_begin] += est[j, :, lo - a:hi - a]
        return out / plan.ratio

    def stitch_concat(self, est, out, start):
        out[:, start:start + est.shape[-1]] = est
