"""Batched form of the streaming per-chunk step, TargetDiarizationStream.asr_audio_streaming
(TargetDiarizationStream.py:189-258), for S concurrent streams (BASELINE.json config 5: 600 ms chunks at batch 256).

The reference serves one stream per object and, per chunk, runs up to one separation and three embeddings one after
the other (TargetASR.multi_speakers_separate_asr, TargetASR.py:571-655, then get_speaker_embedding + is_same_person).
Here the per-stream control flow is kept exactly - same early returns, same state updates, same decisions - but every
GPU-side piece is ONE batched call over all the streams that reach it:

    loudness of the chunks          -> SeparationScoringStage.meter_loudness_device        (AudioProcessor.py:1123-1127)
    enrolment embeddings            -> Embedder.embed_many                                 (TargetASR.py:155-163)
    separation of overlapped chunks -> Separator (one batch per chunk length)              (AudioProcessor.py:885-956)
    embeddings of both streams      -> Embedder.embed_many, cosine, threshold, strict pick (TargetASR.py:612-625)
    segment embedding + decision    -> reused from the batch above / one more embed_many   (TargetASR.py:491-505)

What the reference delegates to other models stays a callable supplied by the caller: `asr(audio, prompt) -> text`
(ASRProcessor.asr_detection), `vad(audio) -> [[start_s, end_s], ...]` (ASRProcessor.vad_detection) and
`audio_preprocess(audio) -> audio` (TargetDiarization.audio_preprocess in stream mode).  Per-stream mutable state
(TargetDiarizationStream.py:23-29) lives in StreamState objects owned by the caller, as SURVEY.md section 8b asks.
"""
import re
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import plan as P


@dataclass
class StreamState:
    """The per-stream fields of TargetDiarizationStream (TargetDiarizationStream.py:23-29) this step reads / writes."""
    current_time: float = 0.0
    target_embedding: Optional[np.ndarray] = None
    prev_asr_text: str = ""
    system_loudness_diff: float = 0.0


def remove_punc(detect_text):
    """TargetDiarizationStream.asr_audio_streaming.remove_punc (:190-196)."""
    if not detect_text:
        return detect_text
    return re.sub(r"[^\w\s]", "", detect_text).lower().strip()


class StageEngine:
    """The three batched compute calls of the step on a SeparationScoringStage (the CPU tests of the control flow
    substitute toy functions with the same signatures)."""

    def __init__(self, stage):
        self.stage = stage

    def meter_many(self, audios):
        """AudioProcessor.meter_loudness of each clip (BS.1770, rounded to 0.1); one device call per clip length."""
        import torch
        st = self.stage
        out = [None] * len(audios)
        by_len = {}
        for i, a in enumerate(audios):
            by_len.setdefault(int(a.shape[0]), []).append(i)
        for _, idx in sorted(by_len.items()):
            x = torch.from_numpy(np.stack([np.asarray(audios[i], dtype=np.float32) for i in idx])).to(st.device)
            for i, v in zip(idx, st.meter_loudness_device(x)):
                out[i] = v
        return out

    def embed_many(self, audios):
        """TargetASR.get_speaker_embedding of each clip -> np.float32 [n, 192]."""
        if not audios:
            return np.zeros((0, 192), np.float32)
        return self.stage.embedder.embed_many([np.asarray(a, dtype=np.float32) for a in audios]).cpu().numpy()

    def separate_many(self, audios):
        """AudioProcessor.separate_speaker of each clip -> [(spk1, spk2)], louder stream first; clips of one length
        share one batched separator call and one batched loudness call."""
        import torch
        st = self.stage
        out = [None] * len(audios)
        by_len = {}
        for i, a in enumerate(audios):
            by_len.setdefault(int(a.shape[0]), []).append(i)
        for T, idx in sorted(by_len.items()):
            x = torch.from_numpy(np.stack([np.asarray(audios[i], dtype=np.float32) for i in idx])).to(st.device)
            est = st.kern.separate(x)                                           # [n, 2, T]
            loud = st.meter_loudness_device(est.reshape(2 * len(idx), T))        # spk1, spk2 of clip 0, of clip 1, ...
            host = st.kern.to_host(est)
            for k, i in enumerate(idx):
                s1, s2 = host[k, 0], host[k, 1]
                if loud[2 * k] < loud[2 * k + 1]:                                # AudioProcessor.py:949-952
                    s1, s2 = s2, s1
                out[i] = (s1, s2)
        return out


def cosine_similarity(a, b):
    """TargetASR.cosine_similarity (TargetASR.py:144-152) on host vectors."""
    a, b = np.asarray(a), np.asarray(b)
    if np.all(a == 0.0) or np.all(b == 0.0):
        return 1.0
    s = np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
    return float(max(0.0, min(s, 1.0)))


def asr_audio_streaming_batch(engine, chunks, states, is_overlap, asr, vad, audio_preprocess=None, vad_inner=None,
                              similarity_threshold=0.4, loudness_diff_threshold=12.0, use_asr_prompt=False,
                              separation_threshold=0.4, is_output_audio=False, update_prev_text=True):
    """One asr_audio_streaming call per stream i on chunks[i] (np.float32 at 16 kHz) with state states[i] and overlap
    flag is_overlap[i] (from the caller's overlap detector, TargetDiarizationStream.py:178-183).  Returns one result
    per stream - None, or the reference's dict {speaker, timerange, text, type, audio} - and updates the states in
    place exactly as S sequential reference calls would (streams do not interact).  vad_inner is the VAD form
    multi_speakers_separate_asr uses on the chunk and on both separated streams (vad_detection with
    min_silence_sec=0.0, TargetASR.py:577-578; default: the same callable as `vad`).  separation_threshold is the
    `threshold` default of multi_speakers_separate_asr (TargetASR.py:571); update_prev_text applies the
    `self.prev_asr_text = result['text']` of process_single_chunk (:184-186)."""
    S = len(chunks)
    if not (len(states) == S and len(is_overlap) == S):
        raise ValueError("one state and one overlap flag per stream")
    prep = audio_preprocess if audio_preprocess is not None else (lambda a: a)
    vad_inner = vad if vad_inner is None else vad_inner
    results = [None] * S
    audio = [np.asarray(c, dtype=np.float32).reshape(-1) for c in chunks]
    overlap = [bool(v) for v in is_overlap]
    # ---- duration gate, clock, prompt (:198-211)
    alive = []
    prompts = [""] * S
    for i in range(S):
        duration = round(audio[i].shape[0] / 16000, 3)
        if duration < 0.4:
            continue
        states[i].current_time = states[i].current_time + duration
        if use_asr_prompt and states[i].prev_asr_text:
            prompts[i] = states[i].prev_asr_text
        alive.append(i)
    # ---- first chunk of a stream without target: it becomes the enrolment sample (:212-217)
    enrol = [i for i in alive if states[i].target_embedding is None]
    if enrol:
        for i, l in zip(enrol, engine.meter_many([audio[i] for i in enrol])):
            states[i].system_loudness_diff = l + 23.0
    for i in alive:
        audio[i] = np.asarray(prep(audio[i]), dtype=np.float32).reshape(-1)
    if enrol:
        emb = engine.embed_many([audio[i] for i in enrol])
        for k, i in enumerate(enrol):
            states[i].target_embedding = emb[k]
            overlap[i] = False
    # ---- loudness gate and VAD (:220-225)
    if alive:
        loud = engine.meter_many([audio[i] for i in alive])
        alive = [i for i, l in zip(alive, loud)
                 if not l < -23.0 + states[i].system_loudness_diff - loudness_diff_threshold]
    vads = {}
    kept = []
    for i in alive:
        v = vad(audio[i])
        if v:
            vads[i] = v
            kept.append(i)
    alive = kept
    # ---- overlapped chunks: separation + scoring of both streams (multi_speakers_separate_asr, TargetASR.py:571-655)
    clips = {i: [] for i in alive}
    seg_emb = {}
    multi = [i for i in alive if overlap[i] and vad_inner(audio[i])]   # its own VAD pass (:594-596): empty -> no clips
    if multi:
        pairs = engine.separate_many([audio[i] for i in multi])
        embs = engine.embed_many([s for pair in pairs for s in pair])
        for k, i in enumerate(multi):
            s1, s2 = pairs[k]
            e1, e2 = embs[2 * k], embs[2 * k + 1]
            sc1 = cosine_similarity(e1, states[i].target_embedding)
            sc2 = cosine_similarity(e2, states[i].target_embedding)
            pick = P.pick_target(sc1, sc2, separation_threshold)
            if pick is None:
                continue
            order = ((s1, e1, sc1), (s2, e2, sc2)) if pick == 1 else ((s2, e2, sc2), (s1, e1, sc1))
            for a, e, sc in order:     # target first, then the other stream; a stream without speech is dropped
                text = asr(a, prompts[i])
                v = vad_inner(a)
                if v:
                    clips[i].append({"timerange": [v[0][0], v[-1][1]], "text": text, "score": round(sc, 2),
                                     "sampling_rate": 16000, "audio": a, "_emb": e})
    for i in alive:
        if not overlap[i]:             # single_speaker_asr (TargetASR.py:658-685)
            clips[i].append({"timerange": [0.0, round(audio[i].shape[0] / 16000, 2)], "text": asr(audio[i], prompts[i]),
                             "score": 1.0, "sampling_rate": 16000, "audio": np.array([], dtype=np.float32)})
    # ---- longest text wins; its audio is embedded and compared with the target (:236-249)
    chosen = {}
    for i in alive:
        lst = clips[i]
        if not lst:
            continue
        if len(lst) > 1:
            lst = sorted(lst, key=lambda x: len(remove_punc(x["text"])), reverse=True)
        text = lst[0]["text"].strip()
        if not text:
            continue
        chosen[i] = (lst[0], text)
    need = [i for i in chosen if not overlap[i]]
    if need:
        emb = engine.embed_many([audio[i] for i in need])
        for k, i in enumerate(need):
            seg_emb[i] = emb[k]
    for i, (clip, text) in chosen.items():
        seg_audio = clip["audio"] if overlap[i] else audio[i]
        e = clip["_emb"] if overlap[i] else seg_emb[i]    # same clip -> same embedding as the reference's second call
        mean = np.mean([e], axis=0)
        is_target = P.is_same_person(cosine_similarity(mean, states[i].target_embedding), similarity_threshold)
        v = vads[i]
        results[i] = {"speaker": "1" if is_target else "0",
                      "timerange": [states[i].current_time + v[0][0], states[i].current_time + v[-1][-1]],
                      "text": text, "type": "overlap" if overlap[i] else "single",
                      "audio": seg_audio if is_output_audio else None}
        if update_prev_text:
            states[i].prev_asr_text = text
    return results
