"""The target-speaker separation + scoring stage, host side (PyTorch for device memory / streams / NCCL only).

  SeparationScoringStage.separate_speaker     <-> AudioProcessor.separate_speaker            (AudioProcessor.py:885-956)
  SeparationScoringStage.wav_chunk_inference  <-> look2hear.utils.wav_chunk_inference       (look2hear/utils/separator.py:72-132)
  SeparationScoringStage.get_speaker_embedding<-> TargetASR.get_speaker_embedding            (TargetASR.py:155-163)
  SeparationScoringStage.cosine_similarity    <-> TargetASR.cosine_similarity                (TargetASR.py:144-152)
  SeparationScoringStage.get_target_embedding <-> TargetASR.get_target_embedding             (TargetASR.py:166-258)
  SeparationScoringStage.same_speaker_batch   <-> TargetDiarizationStream rule 4, batched    (TargetDiarizationStream.py:156-168)
  SeparationScoringStage.separate_and_score   <-> the core of multi_speakers_separate_asr    (TargetASR.py:609-625)
  SeparationScoringStage.score_segments       <-> the per-segment loops                      (TargetDiarization.py:581-629)

Every number is produced by libtdz.so kernels; this file only plans chunks (plan.py), moves buffers and, with
more than one rank, gathers the per-rank spans (one collective at the end)."""
import numpy as np
import torch

from . import plan as P
from .embedder import Embedder
from .separator import Separator


class CudaKernels:
    """The libtdz.so entry points the stage needs, on device tensors.  (The gloo/CPU tests of the sharding logic
    substitute a stand-in for this object; the product has no other implementation.)"""

    def __init__(self, separator, max_workspace_bytes=None):
        self.sep = separator
        self.device = separator.device
        # None: sized from the memory that is free when a batch is planned (a shared or smaller GPU gets smaller
        # sub-batches instead of an out-of-memory error); capped at 100 GiB
        self._max_workspace_bytes = None if max_workspace_bytes is None else int(max_workspace_bytes)

    @property
    def max_workspace_bytes(self):
        if self._max_workspace_bytes is not None:
            return self._max_workspace_bytes
        from . import _lib
        held = self.sep._ws_raw.numel() if self.sep._ws is not None else 0
        if torch.cuda.is_current_stream_capturing():   # no driver queries while a graph is being captured
            return max(held - 1024, 0)
        return min(100 << 30, max((_lib.free_device_bytes(self.device, held) * 3) // 4, held - 1024))

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def to_device(self, audio):
        if isinstance(audio, np.ndarray):
            audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        return audio.to(self.device, torch.float32, non_blocking=True).contiguous()

    def to_host(self, dev):
        """Device tensor -> numpy through page-locked memory (torch's caching host allocator keeps the buffer for
        the next call; a pageable `.cpu()` of an hour of audio costs more than the separation of a window)."""
        host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.device)

    def max_batch(self, T, n=None):
        """Largest sub-batch the workspace budget allows.  With `n` (the batch about to run): if the workspace that is
        ALREADY held fits it, that is the answer - the budget is derived from a driver query of the free memory
        (cudaMemGetInfo: milliseconds, and it waits for the device), which must not sit on the steady-state path."""
        frames = -(-max(T // 8, 1) // 256) * 256
        cap = max(1, min(4096, (1 << 23) // frames))
        if n is not None and self.sep._ws is not None:
            want = min(int(n), cap)
            if self.sep.workspace_bytes(want, T) <= self.sep._ws.numel():
                return want
            hit = self.__dict__.get("_plan")        # the decision taken for this length with this workspace
            if hit is not None and hit[0] == (T, self.sep._ws_generation):
                return hit[1]
        per1 = self.sep.workspace_bytes(1, T)
        nb = int(max(1, min(cap, self.max_workspace_bytes // max(per1, 1))))
        if n is not None:
            # the workspace grows to nb items in the call that follows (generation + 1 if it has to be reallocated)
            grows = self.sep._ws is None or self.sep.workspace_bytes(min(int(n), nb), T) > self.sep._ws.numel()
            self._plan = ((T, self.sep._ws_generation + (1 if grows else 0)), nb)
        return nb

    def separate(self, chunks, out=None, out_strides=None):
        """[n,T] -> [n,2,T] (or into `out` with `out_strides`, see Separator.__call__), in sub-batches that fit the
        workspace budget."""
        n, T = chunks.shape
        nb = self.max_batch(T, n)
        if out is None:
            out = self.empty(n, 2, T)
            out_strides = (2 * T, T)
        elif out_strides is None:
            out_strides = (2 * T, T)
        for i in range(0, n, nb):
            o = out.reshape(-1)[i * out_strides[0]:]
            self.sep(chunks[i:i + nb], out=o, out_strides=out_strides)
        return out

    def gather_segments(self, mix, plan, seg_lo, n_seg, mix_origin=0):
        """Segments [seg_lo, seg_lo + n_seg) of the plan; `mix` holds samples [mix_origin, mix_origin + len(mix))."""
        h = self.sep._h
        seg = self.empty(n_seg, plan.session)
        h.check(h.lib.tdz_gather_segments_span(h.ptr, mix.data_ptr(), mix_origin, mix.shape[0], plan.length,
                                               plan.session, plan.hop, seg_lo, n_seg, seg.data_ptr(), self._stream()),
                "tdz_gather_segments_span")
        return seg

    def stitch_ola(self, est, plan, seg_lo, out_begin, n_out, out=None):
        h = self.sep._h
        if out is None:
            out = self.empty(2, n_out)
        h.check(h.lib.tdz_stitch_ola(h.ptr, est.data_ptr(), plan.session, plan.hop, seg_lo, est.shape[0], plan.length,
                                     out_begin, n_out, float(plan.ratio), out.data_ptr(), self._stream()),
                "tdz_stitch_ola")
        return out


# ------------------------------------------------------------------------------------------------ span engines
def _span_buffer(kern, n, flat):
    """[2, n] output of a span engine: a view of the caller's flat buffer (the rank's slot of the gather) or new."""
    if flat is None:
        return kern.empty(2, n)
    return flat[:2 * n].view(2, n)


def concat_span(kern, mix, bounds, span_begin, span_end, mix_origin=0, out_flat=None):
    """Concat mode over the windows `bounds` (absolute sample indices; `mix` holds the samples from `mix_origin` on).
    Runs of adjacent equal-length windows are ONE batched separator call that reads the windows in place (a view, no
    copy) and writes both streams straight into the stitched output (strided store, no stitch pass).
    Returns [2, span_end - span_begin]."""
    n_out = span_end - span_begin
    out = _span_buffer(kern, n_out, out_flat)
    i = 0
    while i < len(bounds):
        b, e = bounds[i]
        T = e - b
        j = i + 1
        while j < len(bounds) and bounds[j][0] == bounds[j - 1][1] and bounds[j][1] - bounds[j][0] == T:
            j += 1
        n = j - i
        chunks = mix[b - mix_origin:b - mix_origin + n * T].view(n, T)
        # stream s of window k -> out[s, b - span_begin + k T : ...] = flat offset s * n_out + (b - span_begin) + k T
        kern.separate(chunks, out=out.reshape(-1)[b - span_begin:], out_strides=(T, n_out))
        i = j
    return out


def ola_span(kern, mix, plan, out_begin, out_end, seg_lo, seg_hi, batch_size=None, mix_origin=0, out_flat=None):
    """Overlap-add mode for output samples [out_begin, out_end) from segments [seg_lo, seg_hi).  Returns [2, n].
    batch_size bounds the number of segments per separator call (None: as many as the workspace budget allows)."""
    n_seg = seg_hi - seg_lo
    n_out = max(out_end - out_begin, 0)
    out = _span_buffer(kern, n_out, out_flat)
    if n_seg <= 0 or n_out == 0:
        return out
    seg = kern.gather_segments(mix, plan, seg_lo, n_seg, mix_origin)
    if batch_size is None:
        est = kern.separate(seg)
    else:
        est = kern.empty(n_seg, 2, plan.session)
        for i in range(0, n_seg, int(batch_size)):
            kern.separate(seg[i:i + batch_size], out=est[i:i + batch_size])
    return kern.stitch_ola(est, plan, seg_lo, out_begin, n_out, out=out)


def ola_input_range(plan, seg_lo, seg_hi):
    """Samples of [0, L) that segments [seg_lo, seg_hi) read: the rank's output span plus its halo."""
    if seg_hi <= seg_lo:
        return 0, 0
    lo = max(plan.segment_range(seg_lo)[0], 0)
    hi = min(plan.segment_range(seg_hi - 1)[1], plan.length)
    return lo, max(hi, lo)


def _rank_world(group):
    """Sharding is opt-in: without an explicit process group a call never talks to other ranks, whatever the state of
    torch.distributed (one process per GPU normally means every rank holds DIFFERENT audio)."""
    if group is None:
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(group), dist.get_world_size(group)


def _check_same_input(length, group, device):
    """A sharded call is collective: every rank of `group` must hold the same recording.  Cheap guard against
    stitching spans of different inputs together (or hanging in a size-mismatched collective)."""
    import torch.distributed as dist
    t = torch.tensor([length, -length], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    lo, hi = -int(t[1].item()), int(t[0].item())
    if lo != hi:
        raise RuntimeError(f"sharded separation: ranks hold inputs of different lengths ({lo} .. {hi} samples); pass "
                           "group=None for independent per-rank inputs")


def gather_spans(kern, local_flat, span_lens, group, dst=None):
    """The one data-path collective: every rank contributes its [2, n_r] span, stored flat in
    a buffer of 2 * max(n_r) floats; the consumer(s) get the stitched [2, sum n_r] streams.  dst=None: every rank
    (all_gather); dst=r: only rank r (gather), the others return None."""
    import torch.distributed as dist
    world = len(span_lens)
    n_max = max(max(span_lens), 1)
    rank = dist.get_rank(group)
    if dst is None:
        allbuf = torch.empty(world, 2 * n_max, dtype=local_flat.dtype, device=local_flat.device)
        dist.all_gather_into_tensor(allbuf, local_flat.view(1, -1), group=group)
    else:
        allbuf = torch.empty(world, 2 * n_max, dtype=local_flat.dtype, device=local_flat.device) if rank == dst else None
        dist.gather(local_flat, list(allbuf.unbind(0)) if rank == dst else None,
                    dst=dist.get_global_rank(group, dst) if group is not dist.group.WORLD else dst, group=group)
        if rank != dst:
            return None
    total = sum(span_lens)
    out = torch.empty(2, total, dtype=local_flat.dtype, device=local_flat.device)
    pos = 0
    for r, n in enumerate(span_lens):
        if n:
            out[:, pos:pos + n] = allbuf[r, :2 * n].view(2, n)
        pos += n
    return out


def separate_concat(kern, audio, window=P.WINDOW, vad_frames=None, group=None, dst=None):
    """The chunk loop of separate_speaker on the device.  audio [L] (device tensor or host ndarray) -> [2, L_out].

    vad_frames=None is the reference's no-VAD path (one frame [0, L)).  With frames given (low_gpu_ram mode,
    AudioProcessor.py:902-919) the gaps before each frame are zero filled and the output ends at the last frame.
    group = the ranks that share THIS input (sharding is opt-in): rank r separates its contiguous range of windows,
    uploading only those samples when `audio` is a host array; one collective stitches the spans (see gather_spans)."""
    L = int(audio.shape[0])
    rank, world = _rank_world(group)
    if vad_frames is None:
        if world == 1:
            return concat_span(kern, kern.to_device(audio), P.chunk_bounds(L, window), 0, L)
        _check_same_input(L, group, kern.device)
        spans = [P.concat_shard(L, r, world, window) for r in range(world)]
        mine, b, e = spans[rank]
        lens = [s[2] - s[1] for s in spans]
        flat = kern.empty(2 * max(max(lens), 1))
        concat_span(kern, kern.to_device(audio[b:e]), mine, b, e, mix_origin=b, out_flat=flat)
        return gather_spans(kern, flat, lens, group, dst)
    audio = kern.to_device(audio)
    pieces = []
    total = 0
    for i, (fb, fe) in enumerate(vad_frames):
        if fb > total:
            gap = fb if i == 0 else fb - vad_frames[i - 1][1]
            pieces.append(kern.zeros(2, gap))
            total += gap
        pieces.append(concat_span(kern, audio, P.chunk_bounds(fe - fb, window, fb), fb, fe))
        total += fe - fb
    return torch.cat(pieces, dim=1) if pieces else kern.zeros(2, 0)


def separate_ola(kern, audio, sr=16000, target_length=12.0, hop_length=4.0, batch_size=None, group=None, dst=None):
    """wav_chunk_inference on the device.  audio [L] (device tensor or host ndarray) -> [2, L]."""
    L = int(audio.shape[0])
    plan = P.ola_plan(L, sr, target_length, hop_length)
    rank, world = _rank_world(group)
    if world == 1:
        return ola_span(kern, kern.to_device(audio), plan, 0, L, 0, plan.num_session, batch_size)
    _check_same_input(L, group, kern.device)
    shards = [P.ola_shard(plan, r, world) for r in range(world)]
    ob, oe, lo, hi = shards[rank]
    lens = [s[1] - s[0] for s in shards]
    flat = kern.empty(2 * max(max(lens), 1))
    a, b = ola_input_range(plan, lo, hi)      # the span and the halo the rank's segments reach into
    ola_span(kern, kern.to_device(audio[a:b]), plan, ob, oe, lo, hi, batch_size, mix_origin=a, out_flat=flat)
    return gather_spans(kern, flat, lens, group, dst)


# ------------------------------------------------------------------------------------------------ loudness
def _k_weighting(rate):
    """The two K-weighting biquads as pyloudnorm builds them (high shelf +4 dB / 1500 Hz / Q 1/sqrt2, high pass
    38 Hz / Q 0.5; RBJ forms normalised by a0): [(b, a), (b, a)] in float64."""
    out = []
    for G, Q, fc, shelf in ((4.0, 1.0 / np.sqrt(2.0), 1500.0, True), (0.0, 0.5, 38.0, False)):
        A = 10 ** (G / 40.0)
        w0 = 2.0 * np.pi * (fc / rate)
        al, c = np.sin(w0) / (2.0 * Q), np.cos(w0)
        if shelf:
            sq = 2 * np.sqrt(A) * al
            b = [A * ((A + 1) + (A - 1) * c + sq), -2 * A * ((A - 1) + (A + 1) * c), A * ((A + 1) + (A - 1) * c - sq)]
            a = [(A + 1) - (A - 1) * c + sq, 2 * ((A - 1) - (A + 1) * c), (A + 1) - (A - 1) * c - sq]
        else:
            b = [(1 + c) / 2, -(1 + c), (1 + c) / 2]
            a = [1 + al, -2 * c, 1 - al]
        out.append((np.array(b) / a[0], np.array(a) / a[0]))
    return out


def _block_bounds(n_samples, rate, block=0.4):
    """[lo, hi) of the 400 ms / 75 %-overlap blocks with pyloudnorm's float arithmetic (so the integers agree)."""
    T = n_samples / rate
    nblk = int(np.round((T - block) / (block * 0.25)) + 1)
    j = np.arange(nblk, dtype=np.float64)
    # the same float64 operations, in the same order, as pyloudnorm's per-block Python expressions
    # int(block * (j * 0.25) * rate) and int(block * (j * 0.25 + 1) * rate), vectorised (an hour has 36 000 blocks)
    lo = ((block * (j * 0.25)) * rate).astype(np.int64)
    hi = np.minimum(((block * (j * 0.25 + 1)) * rate).astype(np.int64), n_samples)
    return lo, hi


def _gate(z):
    """The two gates of BS.1770 over the block mean squares (absolute -70 LUFS, relative -10 LU) -> LUFS."""
    with np.errstate(divide="ignore", invalid="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
        keep = l >= -70.0
        zg = z[keep].mean() if keep.any() else np.nan
        gamma_r = -0.691 + 10.0 * np.log10(zg) - 10.0
        keep = (l > gamma_r) & (l > -70.0)
        zg = z[keep].mean() if keep.any() else 0.0
        return float(-0.691 + 10.0 * np.log10(zg))


def meter_loudness(audio, rate=16000):
    """AudioProcessor.meter_loudness (AudioProcessor.py:1123-1127): BS.1770 integrated loudness, rounded to 0.1.
    Host numpy like the reference's pyloudnorm call (absent here: restated).  For streams that already live on the
    device use SeparationScoringStage.meter_loudness_device."""
    from scipy.signal import lfilter
    x = np.asarray(audio, dtype=np.float64)
    block = 0.4
    if x.shape[0] < block * rate:
        raise ValueError("Audio must have length greater than the block size.")
    for b, a in _k_weighting(rate):
        x = lfilter(b, a, x)
    lo, hi = _block_bounds(x.shape[0], rate, block)
    cs = np.concatenate(([0.0], np.cumsum(np.square(x))))
    z = (cs[hi] - cs[lo]) / (block * rate)
    return round(_gate(z), 1)


# ------------------------------------------------------------------------------------------------ the stage
def _default_hdbscan_labels(emb):
    """Cluster labels (-1 = noise) as the reference computes them (TargetASR.py:241-242)."""
    try:
        import hdbscan  # the reference's dependency
        return hdbscan.HDBSCAN(min_cluster_size=2, metric="euclidean").fit_predict(emb)
    except ImportError:
        from sklearn.cluster import HDBSCAN
        return HDBSCAN(min_cluster_size=2, metric="euclidean").fit_predict(emb)


class SeparationScoringStage:
    """Owns one Separator and one Embedder on one GPU (one process per GPU; `group` = the ranks that share a long
    input).  Method names and arguments follow the reference methods they stand behind."""

    def __init__(self, separator, embedder, group=None, similarity_threshold=0.0, gather_dst=None):
        """group: the torch.distributed process group whose ranks SHARE every input given to separate_speaker /
        wav_chunk_inference / score_segments (a long recording cut into per-rank spans).  None (default) = this
        stage never communicates: the normal one-process-per-GPU deployment where every rank serves its own audio.
        gather_dst: group rank that receives the stitched result (None = every rank)."""
        self.separator = separator
        self.embedder = embedder
        self.device = separator.device
        self.group = group
        self.gather_dst = gather_dst
        self.kern = CudaKernels(separator)
        from . import _lib
        self._guard = _lib.CallGuard(self.device)    # the static buffers of the captured run() graphs
        self.similarity_threshold = similarity_threshold
        self.is_separate_audio = True

    @classmethod
    def random_init(cls, device="cuda:0", seed=0, **kw):
        """Random-init weights of the named architectures (no checkpoints are shipped with the reference)."""
        from . import synth
        return cls.from_state_dicts(synth.random_state_dict(seed=seed), synth.random_eres2netv2_state_dict(seed=seed),
                                    device, **kw)

    @classmethod
    def from_state_dicts(cls, separator_sd, embedder_sd, device="cuda:0", **kw):
        return cls(Separator(separator_sd, device), Embedder(embedder_sd, device), **kw)

    # ---- AudioProcessor.separate_speaker
    def separate_speaker(self, audio_data, sampling_rate=16000, low_gpu_ram=False, mode="concat", vad_frames=None,
                         resample=None, loudness="device", return_device=False, **ola_kw):
        """np.float32 [L] -> (spk1, spk2) np.float32 [L], louder stream first.

        mode="concat" is the reference rule (10 s windows, bit-exact boundaries); mode="ola" stitches 12 s / 4 s-hop
        segments by overlap-add (wav_chunk_inference).  loudness: "device" (BS.1770 meter on the GPU), a callable
        (audio, rate) -> LUFS such as AudioProcessor.meter_loudness, or None (keep the separator's order).  low_gpu_ram=True uses 1 s windows inside `vad_frames`
        (which the caller's VAD supplies; the reference runs silero-vad there).  `resample(audio, orig_sr,
        target_sr) -> audio` is needed only when sampling_rate != 16000 (the reference calls librosa).

        With a process group (see __init__) the call is collective: every rank passes the same recording, uploads and
        separates only its span (+ overlap-add halo), and the spans are gathered by one NCCL collective; ranks other
        than `gather_dst` return (None, None)."""
        if not self.is_separate_audio:
            return audio_data, audio_data
        orig_sr = sampling_rate
        if sampling_rate != 16000:
            if resample is None:
                raise ValueError("separate_speaker: input is not 16 kHz and no `resample` callable was given")
            audio_data = resample(audio_data, sampling_rate, 16000)
            sampling_rate = 16000
        if low_gpu_ram and vad_frames is None:
            raise ValueError("separate_speaker: low_gpu_ram=True needs `vad_frames` ([[start, end], ...] from the "
                             "caller's VAD, AudioProcessor.py:902-905)")
        window = P.WINDOW_LOW_RAM if low_gpu_ram else P.WINDOW
        if isinstance(audio_data, np.ndarray):
            mix = np.ascontiguousarray(audio_data, dtype=np.float32).reshape(-1)   # spans are uploaded by the engines
        else:
            mix = self.kern.to_device(audio_data).reshape(-1)
        if mode == "concat":
            est = separate_concat(self.kern, mix, window, vad_frames if low_gpu_ram else None, self.group,
                                  self.gather_dst)
        elif mode == "ola":
            est = separate_ola(self.kern, mix, sampling_rate, group=self.group, dst=self.gather_dst, **ola_kw)
        else:
            raise ValueError(f"unknown mode {mode!r}")
        if est is None:        # sharded call, this rank is not the consumer
            return None, None
        if return_device and loudness is None:
            return est[0], est[1]
        swap = False
        if isinstance(loudness, str):  # "device": BS.1770 meter on the GPU, before the streams leave it
            if loudness != "device":
                raise ValueError(f"unknown loudness meter {loudness!r}")
            l1, l2 = self.meter_loudness_device(est, sampling_rate)
            swap = l1 < l2
        if return_device:
            return (est[1], est[0]) if swap else (est[0], est[1])
        host = self.kern.to_host(est)
        spk1, spk2 = host[0], host[1]
        if callable(loudness):
            swap = loudness(spk1, sampling_rate) < loudness(spk2, sampling_rate)
        if swap:
            spk1, spk2 = spk2, spk1
        if orig_sr != sampling_rate:
            spk1, spk2 = resample(spk1, sampling_rate, orig_sr), resample(spk2, sampling_rate, orig_sr)
        return spk1, spk2

    def meter_loudness_device(self, streams, rate=16000):
        """meter_loudness of n device streams [n, L] at once: K-weighting (fp64) and block mean squares on the GPU
        (tdz_loudness_blocks), the two gates over the few thousand block values on the host.  Returns n floats."""
        import ctypes
        x = streams.to(torch.float32).contiguous()
        n, L = x.shape
        block = 0.4
        if L < block * rate:
            raise ValueError("Audio must have length greater than the block size.")
        lo, hi = _block_bounds(L, rate, block)
        coef = np.concatenate([np.concatenate((b, a)) for b, a in _k_weighting(rate)]).astype(np.float64)
        lo_d, hi_d = torch.from_numpy(lo).to(self.device), torch.from_numpy(hi).to(self.device)
        ysq = torch.empty(n, L, dtype=torch.float64, device=self.device)
        z = torch.empty(n, len(lo), dtype=torch.float64, device=self.device)
        h = self.separator._h
        h.check(h.lib.tdz_loudness_blocks(h.ptr, x.data_ptr(), n, L, coef.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                          lo_d.data_ptr(), hi_d.data_ptr(), len(lo), 1.0 / (block * rate),
                                          ysq.data_ptr(), z.data_ptr(), self.kern._stream()), "tdz_loudness_blocks")
        zh = z.cpu().numpy()
        return [round(_gate(zh[i]), 1) for i in range(n)]

    # ---- look2hear.utils.wav_chunk_inference
    def wav_chunk_inference(self, mixture_tensor, sr=16000, target_length=12.0, hop_length=4.0, batch_size=10,
                            n_tracks=2):
        """mixture [1, 1, L] (or [1, L] / [L]) -> [n_tracks=2, 1, L] like the reference with the MossFormer2 adapter
        `lambda x: model(x).unsqueeze(2)`.  batch_size = segments per separator call as in the
        reference (separator.py:115-124); it bounds the scratch memory, not the result (batch-invariant bits).
        batch_size=None lets the workspace budget decide."""
        if n_tracks != 2:
            raise ValueError("the separator has 2 output tracks")
        mix = self.kern.to_device(mixture_tensor).reshape(-1)
        est = separate_ola(self.kern, mix, sr, target_length, hop_length, batch_size, self.group, self.gather_dst)
        return None if est is None else est.unsqueeze(1)

    # ---- TargetASR scoring
    def get_speaker_embedding(self, wav_file, embedding_model="eres2netv2_large"):
        if isinstance(wav_file, np.ndarray):
            wav_file = wav_file.reshape(1, -1)
        return self.embedder(wav_file, output_emb=True)["embs"].reshape(-1)

    @staticmethod
    def cosine_similarity(embedding_a, embedding_b):
        """Scalar form of TargetASR.cosine_similarity for callers that hold host vectors (same rule as the
        tdz_cosine_scores kernel: zero vector -> 1.0, clamp to [0,1])."""
        a = np.asarray(embedding_a)
        b = np.asarray(embedding_b)
        if np.all(a == 0.0) or np.all(b == 0.0):
            return 1.0
        s = np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
        return float(max(0.0, min(s, 1.0)))

    def separate_and_score(self, audio_data, target_embedding, threshold=None, **kw):
        """TargetASR.multi_speakers_separate_asr lines 609-625: separate, embed both streams, cosine vs the target,
        pick.  Returns dict(target=1|2|None, spk1_score, spk2_score, spk1_audio, spk2_audio)."""
        threshold = self.similarity_threshold if threshold is None else threshold
        one = self._separate_and_score_one_window(audio_data, target_embedding, **kw)
        if one is not None:
            spk1, spk2, scores = one
        else:
            spk1, spk2 = self.separate_speaker(audio_data, **kw)
            scores = self.embedder.score_many([spk1, spk2], target_embedding).cpu().tolist()
        return dict(target=P.pick_target(scores[0], scores[1], threshold), spk1_score=scores[0],
                    spk2_score=scores[1], spk1_audio=spk1, spk2_audio=spk2)

    def _separate_and_score_one_window(self, audio_data, target_embedding, sampling_rate=16000, low_gpu_ram=False,
                                       mode="concat", loudness="device", **other):
        """The common case of separate_and_score - a 16 kHz recording the chunk rule keeps as ONE window (up to 15 s) -
        as one upload, one run() (separation, fbank, ERes2NetV2 and cosine of both streams: a single CUDA-graph replay
        from the second call of a length on) and one download, instead of a download of the streams followed by their
        upload for scoring.  Same kernels on the same data as the general path, so the same bits.  Returns None when
        the call is not of that kind (the general path then also produces the reference's errors)."""
        if (other or not self.is_separate_audio or self.group is not None or sampling_rate != 16000 or low_gpu_ram
                or mode != "concat" or not (loudness is None or loudness == "device" or callable(loudness))):
            return None
        if not (isinstance(audio_data, np.ndarray) or torch.is_tensor(audio_data)):
            return None
        L = int(audio_data.size if isinstance(audio_data, np.ndarray) else audio_data.numel())
        if L < 6400 or len(P.chunk_bounds(L, P.WINDOW)) != 1:     # (under 0.4 s the loudness meter refuses the input)
            return None
        mix = self.kern.to_device(audio_data).reshape(1, L)
        est, scores = self.run(mix, target_embedding)
        est = est[0]
        swap = False
        if loudness == "device":
            l1, l2 = self.meter_loudness_device(est, sampling_rate)
            swap = l1 < l2
        host = self.kern.to_host(est)
        sc = scores[0].cpu().tolist()
        spk1, spk2 = host[0], host[1]
        if callable(loudness):
            swap = loudness(spk1, sampling_rate) < loudness(spk2, sampling_rate)
        if swap:
            spk1, spk2, sc = spk2, spk1, [sc[1], sc[0]]
        return spk1, spk2, sc

    def score_segments(self, segments, target_embedding):
        """Batched form of the per-segment loops (TargetDiarization.py:581-629): [N] cosine scores on the device
        (0.0 where the reference's embedding is NaN, i.e. clips of fewer than 9 fbank frames: TargetASR.py:151 computes
        max(0.0, min(nan, 1.0)) = 0.0 for them).  segments: [N,T] tensor / ndarray or a ragged list of clips.

        With a process group the N segments are dealt to the ranks (longest first, round robin - every rank gets the
        same mix of lengths), each rank embeds and scores its share, and one all_gather returns all N scores to every
        rank ("[n_seg] fp32 scores")."""
        rank, world = _rank_world(self.group)
        if world == 1:
            return self.embedder.score_many(segments, target_embedding)
        import torch.distributed as dist
        n = len(segments)
        lengths = [int(segments[i].shape[-1]) for i in range(n)]
        owners = P.deal_segments(lengths, world)
        mine = owners[rank]
        n_max = max(max(len(o) for o in owners), 1)
        local = torch.full((n_max,), float("nan"), dtype=torch.float32, device=self.device)
        if mine:
            if isinstance(segments, (list, tuple)):
                share = [segments[i] for i in mine]
            else:
                share = segments[torch.as_tensor(mine, device=segments.device)] if torch.is_tensor(segments) \
                    else segments[np.asarray(mine)]
            local[:len(mine)] = self.embedder.score_many(share, target_embedding)
        allbuf = torch.empty(world, n_max, dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(allbuf, local.view(1, -1), group=self.group)
        scores = torch.empty(n, dtype=torch.float32, device=self.device)
        for r, idx in enumerate(owners):
            if idx:
                scores[torch.as_tensor(idx, device=self.device)] = allbuf[r, :len(idx)]
        return scores

    def separate_and_score_long(self, audio_data, target_embedding, segment_seconds=4.0, threshold=None,
                                mode="concat", sampling_rate=16000, **kw):
        """The whole stage on one long recording (BASELINE.json config 3): separate (sharded over the stage's process
        group when it has one), cut both separated streams into fixed-length segments, score every segment against
        the target (sharded again, scores gathered) and pick target / non-target per segment (TargetASR.py:612-625
        per segment pair).  The streams stay on the device between the two halves.  Returns a dict:
        spk1 / spk2 (np.float32 [L], louder first; None on ranks other than gather_dst), scores np.float32 [2, n_seg],
        target [n_seg] (1, 2 or 0 for "neither reaches the threshold")."""
        threshold = self.similarity_threshold if threshold is None else threshold
        timings = kw.pop("timings", None)          # optional dict: seconds per phase (each phase is synchronised)

        def mark(name, t_prev):
            if timings is None:
                return t_prev
            import time
            torch.cuda.synchronize(self.device)
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_prev
            return now
        import time
        t = time.perf_counter()
        keep_dst, self.gather_dst = self.gather_dst, None        # scoring shards read the gathered streams
        loudness = kw.pop("loudness", "device")
        try:
            s1, s2 = self.separate_speaker(audio_data, sampling_rate, mode=mode, return_device=True, loudness=None, **kw)
        finally:
            self.gather_dst = keep_dst
        t = mark("upload_separate_gather", t)
        if loudness is not None:       # louder stream first (AudioProcessor.py:949-952), metered on the device
            if loudness != "device":
                raise ValueError("separate_and_score_long meters loudness on the device (loudness='device' or None)")
            l1, l2 = self.meter_loudness_device(torch.stack((s1, s2)), sampling_rate)
            if l1 < l2:
                s1, s2 = s2, s1
        t = mark("loudness", t)
        L = int(s1.shape[0])
        seg = int(segment_seconds * 16000)
        n_seg = L // seg
        if n_seg == 0:
            raise ValueError("recording shorter than one scoring segment")
        rank, _ = _rank_world(self.group)
        out = dict(spk1=None, spk2=None)
        side = None
        if keep_dst is None or rank == keep_dst:
            # the consumer's device -> host copy runs on a side stream, under the scoring of the segments
            host = torch.empty(2, L, dtype=torch.float32, pin_memory=True)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                host[0].copy_(s1, non_blocking=True)
                host[1].copy_(s2, non_blocking=True)
            hn = host.numpy()
            out = dict(spk1=hn[0], spk2=hn[1])
        both = torch.cat((s1[:n_seg * seg], s2[:n_seg * seg])).view(2 * n_seg, seg)
        scores = self.score_segments(both, target_embedding).view(2, n_seg)
        sc = scores.cpu().numpy()
        t = mark("score_segments", t)
        if side is not None:
            side.synchronize()
            s1.record_stream(side)
            s2.record_stream(side)
        t = mark("download_tail", t)
        out["scores"] = sc
        out["target"] = np.array([P.pick_target(float(a), float(b), threshold) or 0 for a, b in zip(sc[0], sc[1])],
                                 dtype=np.int64)
        return out

    # ---- enrolment (TargetASR.get_target_embedding, TargetASR.py:166-258)
    def get_target_embedding(self, target_audio, is_preprocess=True, is_cluster=True, audio_input_type="separate",
                             output_embedding_list=True, sampling_rate=16000, vad=None, loudness_control=None,
                             cluster_labels=None):
        """Same arguments and return as the reference method for array input (one ndarray or a list of ndarrays at
        16 kHz; reading / resampling files is the caller's AudioProcessor).  The reference's own models stay
        callables: `vad(audio) -> [[start_s, end_s], ...]` (FSMN-VAD, ASRProcessor.vad_detection) and
        `loudness_control(audio, sr) -> audio` (AudioProcessor.audio_loudness_control); with is_preprocess=True both
        are required.  All selected pieces are embedded by ONE batched embed_many call (grouped by length) instead
        of one model call per piece.  `cluster_labels` defaults to HDBSCAN(min_cluster_size=2, euclidean) from the
        `hdbscan` package if installed, else scikit-learn's."""
        pieces = [np.array(a, dtype=np.float32, copy=True).reshape(-1) for a in
                  (target_audio if isinstance(target_audio, (list, tuple)) else [target_audio])]
        if is_preprocess:
            if vad is None or loudness_control is None:
                raise ValueError("is_preprocess=True needs the reference's vad and loudness_control callables")
            done = []
            for a in pieces:
                spans = vad(a)
                if not spans:
                    continue
                clips = [a[int(s0 * sampling_rate):int(s1 * sampling_rate)] for s0, s1 in spans]
                clips = [c for c in clips if c.size]
                if clips:
                    a = np.concatenate(clips)
                done.append(np.asarray(loudness_control(a, sampling_rate), dtype=np.float32).reshape(-1))
            pieces = done
        if not pieces:
            return np.zeros([192], dtype=np.float32)
        _, picks = P.enrolment_select([p.shape[0] for p in pieces], sampling_rate, audio_input_type)
        merged = None
        wavs = []
        for src, n in picks:
            if src < 0:
                merged = np.concatenate(pieces) if merged is None else merged
                wavs.append(merged[:n])
            else:
                wavs.append(pieces[src][:n])
        embs = self.embedder.embed_many(wavs).cpu().numpy() if wavs else np.zeros((0, 192), np.float32)
        if is_cluster and len(wavs) > 2 and cluster_labels is None:
            cluster_labels = _default_hdbscan_labels
        return P.enrolment_reduce(list(embs), is_cluster, cluster_labels, output_embedding_list)

    # ---- rule 4 of the streaming gate for many concurrent streams
    def same_speaker_batch(self, prev_audio, current_chunk, threshold=0.4, verbose_result=False):
        """TargetDiarizationStream rule 4 (TargetDiarizationStream.py:156-168) for S streams at once: prev_audio[i] =
        the concatenated earlier chunks of stream i, current_chunk[i] = its newest chunk.  The 2 S clips are
        embedded in one batched call; the decision is TargetASR.is_same_person per stream."""
        S = len(prev_audio)
        if S != len(current_chunk):
            raise ValueError("one current chunk per stream")
        if S == 0:
            return []
        embs = self.embedder.embed_many(list(prev_audio) + list(current_chunk)).cpu().numpy()
        return [P.is_same_person(self.cosine_similarity(embs[i], embs[S + i]), threshold, verbose_result)
                for i in range(S)]

    # ---- the benchmark step: B independent chunks, both streams scored
    def run(self, mix_dev, target_embedding):
        """mix_dev [B,T] on the device -> (est [B,2,T], scores [B,2]); target pick = scores[:,0] > scores[:,1].

        Small calls (the streaming shapes: a few 600 ms chunks) are launch-bound - ~660 kernels of a few microseconds:
        from the second call of a shape on, the whole sequence (separation, fbank, ERes2NetV2, cosine) replays as ONE
        captured CUDA graph."""
        B, T = mix_dev.shape
        sep = self.separator
        if (sep.graph_max_frames and B * int(sep._h.lib.tdz_padded_frames(T)) <= sep.graph_max_frames
                and not torch.cuda.is_current_stream_capturing()):
            r = self._run_graphed(mix_dev, target_embedding, B, T)
            if r is not None:
                return r
        est = self.kern.separate(mix_dev)
        scores = self.embedder.score_many(est.view(2 * B, T), target_embedding)
        return est, scores.view(B, 2)

    def _run_graphed(self, mix_dev, target_embedding, B, T):
        with self._guard, self.separator._guard, self.embedder._guard:
            return self._run_graphed_locked(mix_dev, target_embedding, B, T)

    def _run_graphed_locked(self, mix_dev, target_embedding, B, T):
        cache = self.__dict__.setdefault("_run_graphs", {})
        gen = (self.separator._ws_generation, self.embedder._ws_generation)
        ent = cache.get((B, T))
        if ent is None or ent["gen"] != gen:
            if ent is None:                      # first sight of the shape: eager (also sizes both workspaces)
                cache[(B, T)] = dict(gen=None, graph=None)
                if len(cache) > 16:
                    cache.pop(next(iter(cache)))
                return None
            graphs, sep = self.separator.graph_max_frames, self.separator
            s_in = torch.empty(B, T, dtype=torch.float32, device=self.device)
            s_tgt = torch.empty(192, dtype=torch.float32, device=self.device)
            sep.graph_max_frames = 0             # the inner call must launch, not replay its own graph
            try:
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):  # other threads may allocate meanwhile
                    est = self.kern.separate(s_in)
                    scores = self.embedder.score_many(est.view(2 * B, T), s_tgt)
            finally:
                sep.graph_max_frames = graphs
            ent = dict(gen=(self.separator._ws_generation, self.embedder._ws_generation), graph=g, s_in=s_in,
                       s_tgt=s_tgt, est=est, scores=scores)
            cache[(B, T)] = ent
        ent["s_in"].copy_(mix_dev)
        ent["s_tgt"].copy_(self.embedder._to_dev(target_embedding).reshape(-1))
        ent["graph"].replay()
        return ent["est"].clone(), ent["scores"].clone().view(B, 2)

    def embed(self, wav_dev):
        return self.embedder.embed_many(wav_dev)

    def launches_per_run(self, B, T):
        """Kernels of libtdz.so launched by one run() (memsets/memcpys not counted)."""
        from . import fbank
        n_sep = -(-B // self.kern.max_batch(T))
        frames = fbank.num_frames(T)
        n_emb = -(-2 * B // self.embedder.max_batch(frames))
        return n_sep * Separator.KERNELS_PER_FORWARD + 2 + n_emb * Embedder.KERNELS_PER_FORWARD + 1
