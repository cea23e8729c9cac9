"""`Embedder`: the callable the reference keeps in `TargetASR.embedding['eres2netv2_large']` (TargetASR.py:102-103)
and calls as `self.embedding[name](wav, output_emb=True)['embs']` (TargetASR.py:155-163): Kaldi fbank(80) with
per-utterance mean normalisation, ERes2NetV2-Large, 192-d embedding.  Compute = libtdz.so (tdz_fbank + tdz_embed);
there is no CPU path.  `embed_many` / `score_many` are the batched forms the per-segment scoring loops of the
reference (TargetDiarization.py:581-629) collapse into."""
import ctypes

import numpy as np
import torch

from . import _lib, fbank
from .weights import PackedEres2NetV2

EMBED_DIM = 192


class Embedder:
    sample_rate = 16000
    # launches of one tdz_embed call (csrc/sv_api.cuh): stem 1; per block conv1 + [subsample] + [shortcut] +
    # 4 x implicit-GEMM 3x3 conv + [AFF: 3 x (cat + 2 convs)] + conv3 = 19 + 26 + 92 + 47; layer3_ds 2 (im2col +
    # GEMM); fuse34/TSTP/Linear (split-K + reduce) 6
    KERNELS_PER_FORWARD = 1 + 19 + 26 + 92 + 2 + 47 + 6

    def __init__(self, state_dict=None, device="cuda:0", max_workspace_bytes=None, handle=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tdz.Embedder runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self.device = _lib.resolve_device(self.device)
        self._h = handle if handle is not None else _lib.Handle(self.device.index)
        self._packed = None
        self._ws = None
        self._ws_generation = 0
        # None: a share of the memory that is free when a batch is sized (capped at 24 GiB: larger sub-batches buy
        # nothing, 128 x 4 s utterances need 9 GB)
        self._max_workspace_bytes = None if max_workspace_bytes is None else int(max_workspace_bytes)
        self._guard = _lib.CallGuard(self.device)    # one call at a time per object, any Python thread / stream
        win, tw = fbank.povey_window(), fbank.twiddles()
        mel, lo, hi = fbank.mel_banks()
        self._tables = [torch.from_numpy(np.ascontiguousarray(a)).to(self.device) for a in (win, tw, mel, lo, hi)]
        self._h.check(self._h.lib.tdz_set_fbank_tables(self._h.ptr, *[t.data_ptr() for t in self._tables]),
                      "tdz_set_fbank_tables")
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def load_state_dict(self, state_dict, strict=True):
        self._packed = PackedEres2NetV2(state_dict, self.device)
        self._h.check(self._h.lib.tdz_set_eres2netv2_weights(self._h.ptr, ctypes.byref(self._packed.table)),
                      "tdz_set_eres2netv2_weights")
        return self

    # ---- pieces
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def fbank(self, wav):
        """wav float32 [N,T] on the device -> [N, frames, 80] mean-normalised log-mel features."""
        wav = wav.to(torch.float32).contiguous()
        N, T = wav.shape
        frames = fbank.num_frames(T)
        if frames < 1:
            raise ValueError(f"utterance of {T} samples is shorter than one 25 ms fbank window")
        feat = torch.empty(N, frames, fbank.NMEL, dtype=torch.float32, device=self.device)
        self._h.check(self._h.lib.tdz_fbank(self._h.ptr, wav.data_ptr(), N, T, feat.data_ptr(), self._stream()),
                      "tdz_fbank")
        return feat

    @property
    def max_workspace_bytes(self):
        if self._max_workspace_bytes is not None:
            return self._max_workspace_bytes
        held = self._ws_raw.numel() if self._ws is not None else 0
        if torch.cuda.is_current_stream_capturing():   # no driver queries while a graph is being captured
            return max(held - 1024, 0)
        return min(24 << 30, max(_lib.free_device_bytes(self.device, held) // 4, held - 1024))

    def max_batch(self, frames, n=None):
        """Largest sub-batch whose workspace fits max_workspace_bytes (at least 1).  With `n` (the batch about to run):
        if the workspace already held fits min(n, 256) utterances that is the answer, without the driver query of the
        free memory behind the default budget (milliseconds per call, and it waits for the device)."""
        lib = self._h.lib
        if n is not None and self._ws is not None:
            want = max(1, min(256, int(n)))
            if int(lib.tdz_embed_workspace_bytes(want, frames)) <= self._ws.numel():
                return want
            hit = self.__dict__.get("_plan")            # the decision taken for this shape with this workspace
            if hit is not None and hit[0] == (frames, self._ws_generation):
                return hit[1]
        budget = self.max_workspace_bytes
        per1 = int(lib.tdz_embed_workspace_bytes(1, frames))
        nb = max(1, min(256, budget // max(per1, 1)))
        while nb > 1 and int(lib.tdz_embed_workspace_bytes(nb, frames)) > budget:
            nb -= 1
        if n is not None:
            need = int(lib.tdz_embed_workspace_bytes(min(int(n), nb), frames))
            grows = self._ws is None or need > self._ws.numel()
            self._plan = ((frames, self._ws_generation + (1 if grows else 0)), nb)
        return nb

    def embed_features(self, feat):
        """feat float32 [N, frames, 80] -> [N,192] (ERes2NetV2 forward)."""
        if self._packed is None:
            raise RuntimeError("Embedder has no weights; call load_state_dict first")
        with self._guard:
            return self._embed_features(feat)

    def _embed_features(self, feat):
        N, frames, _ = feat.shape
        emb = torch.empty(N, EMBED_DIM, dtype=torch.float32, device=self.device)
        nb = self.max_batch(frames, N)
        lib, h = self._h.lib, self._h
        for i in range(0, N, nb):
            n = min(nb, N - i)
            nbytes = int(lib.tdz_embed_workspace_bytes(n, frames))
            if self._ws is None or self._ws.numel() < nbytes:
                self._ws_generation += 1
                self._ws = None
                self._ws_raw = None
                self._ws_raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
                off = (-self._ws_raw.data_ptr()) % 1024   # the C ABI wants 1024 B alignment; the allocator gives 512
                self._ws = self._ws_raw[off:off + nbytes]
            h.check(lib.tdz_embed(h.ptr, feat[i:i + n].data_ptr(), n, frames, emb[i:i + n].data_ptr(),
                                  self._ws.data_ptr(), nbytes, self._stream()), "tdz_embed")
        return emb

    def embed_many(self, wavs):
        """Batched embedding.  wavs: device/host tensor [N,T], ndarray [N,T] or a list of 1-D arrays/tensors of
        possibly different lengths.  Returns a device tensor [N,192].

        Ragged lists are grouped by their number of fbank frames m = 1 + (T-400)//160, not by T: with snip_edges the
        features of a clip depend only on its first 400 + 160 (m-1) samples (torchaudio kaldi.py `_get_strided`), so
        clips of one group are cut to that length and share one batched call - same numbers as the reference's
        one-call-per-clip loops (TargetDiarization.py:588-594,612-618), hundreds instead of thousands of launches.
        Clips that leave a single time step after the network's three stride-2 stages (fewer than 9 frames, i.e.
        under 1 680 samples) come back as NaN rows exactly as from the reference model (TargetASR.get_target_embedding
        drops such rows, :235-236; cosine_similarity scores them 0.0, :151)."""
        if isinstance(wavs, (list, tuple)):
            arrs = [self._to_dev(w).reshape(-1) for w in wavs]
            out = torch.empty(len(arrs), EMBED_DIM, dtype=torch.float32, device=self.device)
            by_frames = {}
            for i, a in enumerate(arrs):
                m = fbank.num_frames(a.numel())
                if m < 1:
                    raise ValueError(f"utterance {i} has {a.numel()} samples, shorter than one 25 ms fbank window")
                by_frames.setdefault(m, []).append(i)
            for m, idx in sorted(by_frames.items()):
                n = fbank.WIN + fbank.SHIFT * (m - 1)
                batch = arrs[idx[0]][:n].unsqueeze(0) if len(idx) == 1 else torch.stack([arrs[i][:n] for i in idx])
                emb = self.embed_features(self.fbank(batch))
                if len(idx) == len(arrs):
                    return emb
                out[idx] = emb
            return out
        w = self._to_dev(wavs)
        if w.ndim == 1:
            w = w.unsqueeze(0)
        return self.embed_features(self.fbank(w))

    def score_many(self, wavs, target_embedding):
        """cosine_similarity (TargetASR.py:144-152 semantics) of every utterance against one target: [N]."""
        return self.cosine_scores(self.embed_many(wavs), target_embedding)

    def cosine_scores(self, emb, target_embedding):
        emb = emb.to(torch.float32).contiguous()
        tgt = self._to_dev(target_embedding).reshape(-1).contiguous()
        N, dim = emb.shape
        scores = torch.empty(N, dtype=torch.float32, device=self.device)
        self._h.check(self._h.lib.tdz_cosine_scores(self._h.ptr, emb.data_ptr(), tgt.data_ptr(), N, dim,
                                                    scores.data_ptr(), self._stream()), "tdz_cosine_scores")
        return scores

    def _to_dev(self, x):
        if isinstance(x, np.ndarray):
            if np.issubdtype(x.dtype, np.integer):  # the modelscope pipeline rescales integer PCM by 2^-15
                x = x.astype(np.float32) / 32768.0
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to(self.device, torch.float32)

    # ---- the reference call surface (TargetASR.py:161)
    def __call__(self, wav, output_emb=True):
        if isinstance(wav, str) or (isinstance(wav, (list, tuple)) and wav and isinstance(wav[0], str)):
            raise TypeError("tdz.Embedder takes waveforms (ndarray [n,T] / list of arrays); file paths are read by "
                            "the caller (AudioProcessor.read_audio)")
        embs = self.embed_many(wav).cpu().numpy()
        out = {"embs": embs}
        if len(embs) == 2:  # the pipeline also scores a pair when given two inputs
            a, b = embs
            out["outputs"] = {"score": float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-12))}
        return out
