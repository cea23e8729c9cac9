"""CPU tests of the batched streaming step (targetdiarization_b200/streaming.py) against golden vectors produced by
running the REFERENCE method TargetDiarizationStream.asr_audio_streaming (and the TargetASR methods it calls) from
source with the toy models of oracle/stream_toys.py (oracle/make_golden.py::make_streaming)."""
import os

import numpy as np
import pytest

from oracle import stream_toys as T
from targetdiarization_b200 import streaming

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(chunks, overlap, engine, batched=True):
    n_steps, n_streams = len(chunks), len(chunks[0])
    states = [streaming.StreamState() for _ in range(n_streams)]
    out = []
    kw = dict(asr=T.asr, vad=lambda a: T.vad(a), vad_inner=lambda a: T.vad(a, 0.0), audio_preprocess=T.audio_preprocess,
              similarity_threshold=0.4, loudness_diff_threshold=12.0, use_asr_prompt=True)
    for s in range(n_steps):
        if batched:
            res = streaming.asr_audio_streaming_batch(engine, chunks[s], states, overlap[s], **kw)
        else:   # one stream per call: what S reference objects would do
            res = [streaming.asr_audio_streaming_batch(engine, [chunks[s][k]], [states[k]], [overlap[s][k]], **kw)[0]
                   for k in range(n_streams)]
        snap = [(st.current_time, st.system_loudness_diff,
                 np.zeros(192, np.float32) if st.target_embedding is None else st.target_embedding.copy())
                for st in states]
        out.append((res, snap))
    return out


def test_batched_step_equals_reference_run():
    gd = np.load(os.path.join(GOLDEN, "streaming.npz"))
    chunks, overlap = T.scenario()
    n_steps, n_streams = (int(v) for v in gd["shape"])
    assert (n_steps, n_streams) == (len(chunks), len(chunks[0]))
    engine = T.ToyEngine()
    got = _run(chunks, overlap, engine)
    texts = list(gd["texts"])
    n_results = n_overlap = 0
    for s in range(n_steps):
        res, snap = got[s]
        for k in range(n_streams):
            want = gd[f"r{s}_{k}"]
            r = res[k]
            assert (r is None) == (want[0] == 0.0), (s, k)
            if r is not None:
                n_results += 1
                n_overlap += r["type"] == "overlap"
                assert float(r["speaker"]) == want[1], (s, k)
                assert r["timerange"] == [want[2], want[3]], (s, k, r["timerange"])      # same float arithmetic
                assert (r["type"] == "overlap") == bool(want[4])
                assert r["text"] == texts[s * n_streams + k] and r["audio"] is None
            ct, sld, te = snap[k]
            w = gd[f"s{s}_{k}"]
            assert ct == w[0] and sld == w[1], (s, k)
            assert float(te.sum()) == w[2] and float(np.abs(te).sum()) == w[3], (s, k)
    assert n_results == 17 and n_overlap >= 3
    # the point of the batched form: per step ONE separation batch and a handful of embedding / loudness batches,
    # not one model call per stream
    assert engine.calls["separate"] <= n_steps and engine.calls["embed"] <= 3 * n_steps
    assert engine.calls["meter"] <= 2 * n_steps


def test_batched_equals_stream_at_a_time():
    chunks, overlap = T.scenario(seed=5, n_streams=6, n_steps=5)
    a = _run(chunks, overlap, T.ToyEngine(), batched=True)
    b = _run(chunks, overlap, T.ToyEngine(), batched=False)
    for (ra, sa), (rb, sb) in zip(a, b):
        assert ra == rb
        for x, y in zip(sa, sb):
            assert x[0] == y[0] and x[1] == y[1] and np.array_equal(x[2], y[2])


def test_argument_checks():
    with pytest.raises(ValueError):
        streaming.asr_audio_streaming_batch(T.ToyEngine(), [np.zeros(9600, np.float32)], [], [False], asr=T.asr, vad=T.vad)
    assert streaming.remove_punc("Ab, c!") == "ab c" and streaming.remove_punc("") == ""
