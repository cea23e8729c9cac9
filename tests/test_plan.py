"""CPU tests of the host-side planner (targetdiarization_b200/plan.py): chunk rule, overlap-add plan, rank shards."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import stage_port
from targetdiarization_b200 import plan


@given(st.integers(1, 5_000_000), st.sampled_from([16000, 160000, 1000]))
@settings(max_examples=300, deadline=None)
def test_chunk_bounds_properties(L, window):
    b = plan.chunk_bounds(L, window)
    assert b == stage_port.chunk_bounds(L, window)
    assert b[0][0] == 0 and b[-1][1] == L
    for (s0, e0), (s1, e1) in zip(b, b[1:]):
        assert e0 == s1
    lens = [e - s for s, e in b]
    assert all(n == window for n in lens[:-1])
    assert 0 < lens[-1] <= 1.5 * window
    if L >= window:
        assert lens[-1] > window / 2


def test_chunk_bounds_edge_cases():
    W = 160000
    assert plan.chunk_bounds(1) == [(0, 1)]
    assert plan.chunk_bounds(W - 1) == [(0, W - 1)]
    assert plan.chunk_bounds(W) == [(0, W)]
    assert plan.chunk_bounds(W + W // 2) == [(0, W + W // 2)]               # remainder == W/2 extends
    assert plan.chunk_bounds(W + W // 2 + 1) == [(0, W), (W, W + W // 2 + 1)]  # remainder > W/2 is its own window
    assert plan.chunk_bounds(57_600_000)[-1] == (57_440_000, 57_600_000) and len(plan.chunk_bounds(57_600_000)) == 360
    assert plan.chunk_bounds(100, 160000, start=40) == [(40, 140)]


@given(st.integers(1, 3_000_000), st.integers(1, 8))
@settings(max_examples=200, deadline=None)
def test_concat_shards_partition_the_windows(L, world):
    bounds = plan.chunk_bounds(L)
    got, pos = [], 0
    for r in range(world):
        mine, b, e = plan.concat_shard(L, r, world)
        assert b == pos
        pos = e
        got += mine
    assert got == bounds and pos == L


@given(st.integers(1, 200_000), st.integers(1, 8))
@settings(max_examples=200, deadline=None)
def test_ola_shards_cover_every_sample_with_all_its_segments(L, world):
    p = plan.ola_plan(L, 1000, 12.0, 4.0)
    assert p.num_session == stage_port.ola_plan(L, 1000)[3]
    pos = 0
    for r in range(world):
        ob, oe, lo, hi = plan.ola_shard(p, r, world)
        assert ob == pos and oe >= ob
        pos = oe
        for n in {ob, (ob + oe) // 2, oe - 1} if oe > ob else ():
            # brute force: segments whose window [i*hop - pad, i*hop - pad + session) contains n
            cover = [i for i in range(p.num_session) if p.segment_range(i)[0] <= n < p.segment_range(i)[1]]
            assert cover and lo <= cover[0] and cover[-1] < hi
            assert len(cover) == 3  # rectangular OLA: every sample is covered exactly ratio = 3 times
    assert pos == L


def test_ola_plan_c3_numbers():
    """1 h at 16 kHz -> 903 segments of 192 000 samples."""
    p = plan.ola_plan(57_600_000)
    assert (p.session, p.hop, p.pad, p.num_session, p.ratio) == (192000, 64000, 128000, 903, 3.0)
    with pytest.raises(ValueError):
        plan.ola_plan(1000, 16000, 4.0, 4.0)


def test_pick_target_rule():
    """TargetASR.py:612-625: strict > picks spk1, ties go to spk2, both under threshold -> nothing."""
    assert plan.pick_target(0.7, 0.3) == 1
    assert plan.pick_target(0.3, 0.7) == 2
    assert plan.pick_target(0.5, 0.5) == 2
    assert plan.pick_target(0.1, 0.15, threshold=0.2) is None
    assert plan.pick_target(0.1, 0.25, threshold=0.2) == 2
    for a, b, t in ((0.7, 0.3, 0.0), (0.5, 0.5, 0.0), (0.1, 0.15, 0.2)):
        assert plan.pick_target(a, b, t) == stage_port.pick_target(a, b, t)


def test_meter_loudness_known_answers():
    """BS.1770: a full-scale 997 Hz sine reads -3.01 LUFS (mono); product (host numpy) == oracle restatement."""
    from targetdiarization_b200 import meter_loudness
    sr = 16000
    t = np.arange(sr * 3) / sr
    x = np.sin(2 * np.pi * 997.0 * t).astype(np.float32)
    assert meter_loudness(x, sr) == pytest.approx(-3.0, abs=0.15)  # 16 kHz RBJ filters: -3.06
    assert meter_loudness(0.1 * x, sr) == pytest.approx(-23.0, abs=0.15)
    rng = np.random.default_rng(1)
    y = (rng.standard_normal(sr * 5) * 0.05 * (np.sin(2 * np.pi * 0.7 * np.arange(sr * 5) / sr) > 0)).astype(np.float32)
    assert meter_loudness(y, sr) == stage_port.meter_loudness(y, sr)
    with pytest.raises(ValueError):
        meter_loudness(x[:6000], sr)   # pyloudnorm raises under one 400 ms block


def test_per_segment_rules_match_reference_source():
    """target_embedding_to_target_spk / recheck_target_speaker / is_same_person: the product's score rules against the
    reference functions themselves (extracted from the reference source with `ast` in the build container) on random
    scores; everywhere: fixed known answers."""
    import random
    from tests.conftest import has_reference
    assert plan.target_spk_from_scores(["a", "b", "a", "c"], [0.2, 0.5, 0.9, 0.55]) == "a"     # means .55 .5 .55: first of the tie
    assert plan.target_spk_from_scores([], []) == ""
    res = [{"speaker": "a", "audio": 1}, {"speaker": "b", "audio": 1}, {"speaker": "a", "audio": None}]
    out = plan.recheck_target_speaker([dict(r) for r in res], [0.1, 0.9, None], "a", 0.2)
    assert [r["speaker"] for r in out] == ["-1", "b", "a"] and [r["score"] for r in out] == [0.1, -1.0, -1.0]
    out = plan.recheck_target_speaker([dict(r) for r in res], [0.1, 0.9, None], "a", 0.2, method="recheck_both")
    assert [r["speaker"] for r in out] == ["-1", "a", "a"]
    assert plan.is_same_person(0.4) is True and plan.is_same_person(0.39999) is False
    assert plan.is_same_person(0.1234567, verbose_result=True) == {"is_same": False, "score": 0.123}
    if not has_reference():
        return
    import types
    from oracle.make_golden import extract_method
    ns = {"np": np, "Literal": __import__("typing").Literal, "Union": __import__("typing").Union}
    code = extract_method("/root/reference/TargetDiarization.py", "TargetDiarization", "recheck_target_speaker")
    exec(compile(code, "ref", "exec"), ns)
    ref_recheck = ns["recheck_target_speaker"]
    rng = random.Random(3)
    for trial in range(50):
        n = rng.randint(1, 8)
        result = [{"speaker": rng.choice(["a", "b", "c"]), "audio": (None if rng.random() < 0.2 else i)} for i in range(n)]
        scores = [rng.random() for _ in range(n)]
        thr = rng.choice([0.0, 0.2, 0.5])
        method = rng.choice(["recheck_target", "recheck_others", "recheck_both"])
        fake = types.SimpleNamespace(
            target_similarity_threshold=thr,
            tasr=types.SimpleNamespace(get_speaker_embedding=lambda wav_file: wav_file,
                                       cosine_similarity=lambda embedding_a, embedding_b: scores[embedding_b]))
        want = ref_recheck(fake, [dict(r) for r in result], "a", "tgt", method)
        got = plan.recheck_target_speaker([dict(r) for r in result],
                                          [None if r["audio"] is None else scores[r["audio"]] for r in result], "a", thr,
                                          method)
        assert got == want, (trial, got, want)


# ---------------------------------------------------------------------------------------------- enrolment rules
def _toy_embed(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    e = np.concatenate(([a.size, a.sum(), a[:3].sum()], np.cos(a.size * np.arange(1, 190) * 1e-3)))
    if a.size % 13 == 0:
        e[1] = np.nan
    return e.astype(np.float32)


@given(st.lists(st.integers(1, 70000), min_size=0, max_size=7),
       st.sampled_from(["auto", "separate", "merge", "longest"]), st.booleans(), st.booleans(),
       st.lists(st.sampled_from([-1, 0, 1]), min_size=7, max_size=7), st.integers(0, 2**31 - 1))
@settings(max_examples=150, deadline=None)
def test_enrolment_rules_equal_the_oracle_restatement(lengths, mode, is_cluster, as_list, labels, seed):
    """plan.enrolment_select + enrolment_reduce == oracle.stage_port.get_target_embedding (the line-by-line
    restatement of TargetASR.py:203-258) for random piece lengths, modes and cluster labels."""
    g = np.random.default_rng(seed)
    pieces = [(g.standard_normal(n) * 0.1).astype(np.float32) for n in lengths]
    lab = lambda e: np.array(labels)[:len(e)]
    want = stage_port.get_target_embedding(pieces, _toy_embed, lab, is_cluster, mode, as_list)
    if not pieces:
        assert np.array_equal(want, np.zeros(192, np.float32))
        return
    _, picks = plan.enrolment_select(lengths, 16000, mode)
    merged = pieces[0] if len(pieces) == 1 else np.concatenate(pieces)
    embs = [_toy_embed((merged if s < 0 else pieces[s])[:n]) for s, n in picks]
    got = plan.enrolment_reduce(embs, is_cluster, lab, as_list)
    if as_list:
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    else:
        assert np.array_equal(np.asarray(got), np.asarray(want))
