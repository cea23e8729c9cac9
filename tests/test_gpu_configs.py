"""GPU parity tests of the BASELINE.json configurations at their own shapes, through the call surfaces the reference
uses (numpy in, numpy out) and against golden vectors produced by the reference module itself
(oracle/make_golden.py) or, where the reference has no source (ERes2NetV2), against the oracle port.

  C1  assets/chat_mix.wav (whole file, 8.665 s) + assets/female_a.wav target   target_diarization_test.py:26-40
  C2  one item of the benchmark batch at T = 64 000
  C4  per-segment scoring: ragged segments vs the port, 4 096 segments as properties  TargetDiarization.py:581-629
  C5  streaming shape: 600 ms chunks at batch 1 and batch 256                  TargetDiarizationStream.py:189-258
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _record(name, value):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


def _snr_db(ref, est):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(est, dtype=np.float64) - ref
    return float(10 * np.log10(np.sum(ref * ref) / max(np.sum(err * err), 1e-300)))


@pytest.fixture(scope="module")
def stage():
    import torch
    from targetdiarization_b200 import SeparationScoringStage
    return torch, SeparationScoringStage.random_init("cuda:0", seed=0)


# ------------------------------------------------------------------------------------------------------------ C1
def test_c1_chat_mix_whole_file_against_reference_run():
    """Config C1 end to end through the drop-in surface: separate_speaker(np) on the whole demo mixture (one 8.665 s
    window: S = 17 328 frames, 68 attention groups, 9 linear-attention splits) against the reference module's own
    output; then get_speaker_embedding of both streams and of female_a.wav, cosine_similarity and the pick, against
    the oracle embedder run on the REFERENCE-separated streams."""
    import torch
    from targetdiarization_b200.synth import random_eres2netv2_state_dict, random_state_dict
    from targetdiarization_b200 import SeparationScoringStage, plan
    gd = np.load(os.path.join(GOLDEN, "c1_chat_mix.npz"))
    st = SeparationScoringStage.from_state_dicts(random_state_dict(seed=0, perturb=True),
                                                 random_eres2netv2_state_dict(seed=0), "cuda:0")
    mix = gd["mix_pcm"].astype(np.float32) / 32768.0          # AudioProcessor.int16_to_float32
    assert mix.shape == (138634,) and plan.chunk_bounds(mix.shape[0]) == [(0, 138634)]
    s1, s2 = st.separate_speaker(mix, loudness=None)
    assert s1.shape == mix.shape and s1.dtype == np.float32
    snr = _snr_db(gd["out_stride4"], np.stack((s1, s2))[:, ::4])
    _record("c1_chat_mix_whole_file_vs_reference_run_snr_db", snr)
    assert snr >= 40.0, snr
    # the last two samples are the zero pad of mossformer2.py:585-586 (T' = 138 632)
    assert not s1[-2:].any() and not s2[-2:].any()
    target = gd["target_pcm"]                                  # int16 PCM: the pipeline rescales by 2^-15
    emb_t = st.get_speaker_embedding(target)
    e1, e2 = st.get_speaker_embedding(s1), st.get_speaker_embedding(s2)
    cos = lambda a, b: float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))
    cs = [cos(e1, gd["emb"][0]), cos(e2, gd["emb"][1]), cos(emb_t, gd["emb_target"])]
    _record("c1_embedding_cosine_vs_oracle_min", min(cs))
    assert min(cs) >= 0.999, cs
    sc = [st.cosine_similarity(e1, emb_t), st.cosine_similarity(e2, emb_t)]
    _record("c1_score_abs_err_max", max(abs(sc[k] - gd["scores"][k]) for k in range(2)))
    assert max(abs(sc[k] - gd["scores"][k]) for k in range(2)) < 5e-4
    if abs(gd["scores"][0] - gd["scores"][1]) > 0.02:          # the pick is only defined up to the score tolerance
        assert (plan.pick_target(sc[0], sc[1], 0.0) or 0) == int(gd["pick"][0])
    # the same through the one-call form
    r = st.separate_and_score(mix, emb_t, loudness=None)
    assert r["spk1_score"] == pytest.approx(sc[0], abs=1e-6) and r["spk2_score"] == pytest.approx(sc[1], abs=1e-6)
    del st
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------ C2
def test_c2_item_at_benchmark_shape_against_reference_run(stage):
    """Item 0 of the benchmark batch (T = 64 000, S = 7 999) against the reference module's output on the same
    mixture and weights; the item is taken out of the FULL batch-of-64 call the benchmark times."""
    torch, st = stage
    from targetdiarization_b200.synth import synthetic_mixture
    gd = np.load(os.path.join(GOLDEN, "c2_item.npz"))
    mix = synthetic_mixture(64, 64000, seed=1234)
    est = st.separator(mix.cuda())
    snr = _snr_db(gd["out_stride4"], est[0, :, ::4].cpu().numpy())
    _record("c2_item_T64000_vs_reference_run_snr_db", snr)
    assert snr >= 40.0, snr
    s1, s2 = st.separate_speaker(mix[0].numpy(), loudness=None)      # the drop-in surface, same bits
    assert np.array_equal(np.stack((s1, s2)), est[0].cpu().numpy())


# ------------------------------------------------------------------------------------------------------------ C4
def _ragged(n, seed, lo=3000, hi=48000):
    from targetdiarization_b200.synth import synthetic_mixture
    g = np.random.default_rng(seed)
    lens = g.integers(lo, hi, size=n)
    base = synthetic_mixture(8, hi, seed=seed).numpy()
    return [np.ascontiguousarray(base[i % 8, int(g.integers(0, hi - m + 1)):][:m] * g.uniform(0.3, 1.0)) for i, m in
            enumerate(lens)]


def test_c4_ragged_segments_against_oracle(stage):
    """Per-segment scoring (TargetDiarization.py:581-629 collapsed into one batched call): 48 ragged clips of
    0.19 - 3 s, among them clips under 1 680 samples (NaN embedding in the reference model, which its cosine turns into a score of 0.0)
    and one clip that is a prefix of another (the frame-count grouping must not mix them up)."""
    torch, st = stage
    from oracle import eres2netv2_port as E
    from oracle import stage_port
    esd = E.random_state_dict(seed=0)
    segs = _ragged(44, seed=3)
    segs += [segs[5][:1679].copy(), segs[6][:500].copy(), segs[7][:1680].copy(), segs[9][:segs[9].shape[0] - 37].copy()]
    target = E.embed(esd, torch.from_numpy(_ragged(1, seed=9, lo=30000, hi=30001)[0])[None])[0].numpy()
    got = st.score_segments(segs, target).cpu().numpy()
    assert got.shape == (48,)
    worst_cos, worst_score = 1.0, 0.0
    ref_all, got_all = [], []
    embs = st.embedder.embed_many(segs).cpu().numpy()
    for i, s in enumerate(segs):
        with torch.no_grad():
            ref = E.embed(esd, torch.from_numpy(s)[None])[0].numpy()
        if np.isnan(ref).any():
            # NaN embedding -> score 0.0: max(0.0, min(nan, 1.0)) in TargetASR.cosine_similarity (TargetASR.py:151)
            assert s.shape[0] < 1680 and np.isnan(embs[i]).all() and got[i] == 0.0
            assert stage_port.cosine_similarity(ref, target) == 0.0
            continue
        assert not np.isnan(embs[i]).any()
        ref_all.append(ref)
        got_all.append(embs[i])
        worst_cos = min(worst_cos, float(np.dot(ref, embs[i]) / (np.linalg.norm(ref) * np.linalg.norm(embs[i]))))
        worst_score = max(worst_score, abs(stage_port.cosine_similarity(ref, target) - float(got[i])))
    _record("c4_ragged_embedding_cosine_min", worst_cos)
    _record("c4_ragged_score_abs_err_max", worst_score)
    assert worst_cos >= 0.999 and worst_score < 5e-4
    # A random-init embedder maps every clip close to one common vector (cosine ~ 0.999 between ANY two clips), so the
    # contract's cosine is a weak check here: also require the error to be small against the spread BETWEEN clips.
    refs = np.stack(ref_all)
    spread = float(np.median(np.linalg.norm(refs[:, None] - refs[None], axis=-1)[np.triu_indices(len(refs), 1)]))
    err = max(float(np.linalg.norm(a - b)) for a, b in zip(got_all, ref_all))
    _record("c4_ragged_embedding_err_over_spread", err / spread)
    assert err / spread < 0.25, (err, spread)
    assert int(np.isnan(embs).any(axis=1).sum()) == 2 and not np.isnan(got).any()
    # the scalar rules on top of the batched scores (TargetDiarization.py:581-600, 603-629)
    from targetdiarization_b200 import plan
    speakers = [str(i % 3) for i in range(48)]
    assert plan.target_spk_from_scores(speakers, got.tolist()) in ("0", "1", "2")


def test_c4_full_size_properties(stage):
    """4 096 segments (BASELINE config 4) in one call: scores in [0, 1]; every score equals the score of the same
    clip embedded on its own batch (batch-invariant bits); permuting the input permutes the output."""
    torch, st = stage
    from targetdiarization_b200.synth import synthetic_mixture
    base = synthetic_mixture(64, 64000 + 4096, seed=41).cuda()
    idx = torch.arange(4096, device="cuda")
    segs = torch.stack([base[i % 64, (i // 64) * 64:(i // 64) * 64 + 64000] for i in range(4096)])   # [4096, 64000]
    tgt = st.embed(synthetic_mixture(1, 48000, seed=42).cuda())[0]
    scores = st.score_segments(segs, tgt)
    assert scores.shape == (4096,) and bool(torch.isfinite(scores).all())
    assert bool(((scores >= 0) & (scores <= 1)).all())
    for i in (0, 777, 4095):
        assert torch.equal(st.score_segments(segs[i:i + 1], tgt)[0], scores[i])
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(st.score_segments(segs[perm[:512]], tgt), scores[perm[:512]])
    del segs, idx


# ------------------------------------------------------------------------------------------------------------ C5
def test_c5_streaming_chunk_shapes_against_oracle(stage):
    """600 ms chunks (T = 9 600, S = 1 199, 5 attention groups): batch 1 and batch 256 against the oracle port, and
    batch 256 == 256 single calls bit for bit (concurrent streams must not influence each other)."""
    torch, st = stage
    from oracle.mossformer2_port import mossformer2_forward, snr_db
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200.synth import synthetic_mixture
    sd = random_state_dict(seed=0)
    mix = synthetic_mixture(256, 9600, seed=51)
    one = st.separator(mix[:1].cuda()).cpu()
    big = st.separator(mix.cuda()).cpu()
    assert big.shape == (256, 2, 9600)
    for i in (0, 100, 255):
        with torch.no_grad():
            ref = mossformer2_forward(sd, mix[i:i + 1])
        snr = snr_db(ref, big[i:i + 1])
        _record(f"c5_batch256_item{i}_snr_db", snr)
        assert snr >= 40.0, (i, snr)
    assert torch.equal(one[0], big[0])
    assert torch.equal(st.separator(mix[100:101].cuda()).cpu()[0], big[100])
    # rule 4 of the streaming gate on the same chunks (two clips per stream, one batched embedding call)
    prev = [mix[i].numpy() for i in range(4)]
    cur = [mix[4 + i].numpy() for i in range(4)]
    res = st.same_speaker_batch(prev, cur, threshold=0.4, verbose_result=True)
    assert len(res) == 4 and all(0.0 <= r["score"] <= 1.0 for r in res)


# ------------------------------------------------------------------------------------------------ checkpoint loading
def test_from_pretrain_round_trip(tmp_path):
    """A file in BaseModel.serialize layout (base_model.py:132-146) loads through Separator.from_pretrain with the
    reference's call form `from_pretrain(path, **cfg.model)` and gives the same bits as the state dict itself;
    another architecture - by config, by model_args or by tensor shapes - is refused."""
    import torch
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    from targetdiarization_b200.separator import MODEL_ARGS
    sd = random_state_dict(seed=5, perturb=True)
    path = str(tmp_path / "best_model.pth")
    torch.save(dict(model_name="MossFormer2", state_dict=sd, model_args=dict(MODEL_ARGS),
                    infos=dict(software_versions=dict(torch_version=torch.__version__))), path)
    cfg_model = dict(MODEL_ARGS)                             # config.yaml `model:` block minus _target_
    a = Separator.from_pretrain(path, **cfg_model)
    b = Separator(sd, "cuda:0")
    x = torch.randn(1, 5000, generator=torch.Generator().manual_seed(0)).cuda() * 0.1
    assert torch.equal(a(x), b(x))
    assert a.eval() is a and a.to("cuda:0") is a
    with pytest.raises(ValueError):
        Separator.from_pretrain(path, **dict(cfg_model, num_blocks=12))
    with pytest.raises(TypeError):
        Separator.from_pretrain(path, hidden=3)
    torch.save(dict(model_name="MossFormer2", state_dict=sd, model_args=dict(MODEL_ARGS, num_spks=3)), path)
    with pytest.raises(ValueError):
        Separator.from_pretrain(path)
    torch.save(dict(model_name="ConvTasNet", state_dict=sd), path)
    with pytest.raises(ValueError):
        Separator.from_pretrain(path)
    small = {k: v for k, v in sd.items() if ".23." not in k}  # a 23-layer checkpoint
    torch.save(dict(model_name="MossFormer2", state_dict=small), path)
    with pytest.raises(RuntimeError):
        Separator.from_pretrain(path)


def test_strided_output_and_device_resolution():
    """tdz_separate_strided writes windows straight into the stitched [2, L] layout; 'cuda' means the current
    device; creating a handle does not change the caller's current device."""
    import torch
    from targetdiarization_b200.synth import random_state_dict
    from targetdiarization_b200 import Separator
    sep = Separator(random_state_dict(seed=0), "cuda")
    assert sep.device == torch.device("cuda", torch.cuda.current_device())
    x = (torch.randn(3, 4000, generator=torch.Generator().manual_seed(1)) * 0.1).cuda()
    ref = sep(x)
    out = torch.full((2, 3 * 4000 + 7), 7.0, device="cuda")
    sep(x, out=out.reshape(-1)[5:], out_strides=(4000, out.shape[1]))
    for k in range(3):
        assert torch.equal(out[:, 5 + 4000 * k:5 + 4000 * (k + 1)], ref[k])
    assert bool((out[:, :5] == 7.0).all()) and bool((out[:, -2:] == 7.0).all())
    with pytest.raises(RuntimeError):
        sep(x, out=out, out_strides=(100, 100))


def test_cuda_graph_replay_equals_eager(stage):
    """Small (launch-bound) calls replay a captured CUDA graph from the second call of a shape on: same kernels, same
    bits as the eager launch sequence, for the separator alone and for the whole run() step; a workspace that moves
    (larger call in between) invalidates the captured graphs instead of leaving them with stale pointers."""
    torch, st = stage
    from targetdiarization_b200.synth import synthetic_mixture
    sep = st.separator
    mix = synthetic_mixture(3, 9600, seed=61).cuda()
    tgt = st.embed(synthetic_mixture(1, 16000, seed=62).cuda())[0]
    keep = sep.graph_max_frames
    sep.graph_max_frames = 0
    eager = sep(mix[:1]).clone()
    eager_run = [t.clone() for t in st.run(mix, tgt)]
    sep.graph_max_frames = keep
    sep._graphs.clear()
    first = sep(mix[:1]).clone()          # eager (first sight of the shape)
    second = sep(mix[:1]).clone()         # captured + replayed
    third = sep(mix[:1] * 1.0).clone()    # replayed
    assert sep._graphs[(1, 9600)][1] is not None and sep._graphs[(1, 9600)][0] >= 3
    assert torch.equal(first, eager) and torch.equal(second, eager) and torch.equal(third, eager)
    other = sep(mix[1:2])                 # different data through the same graph
    sep.graph_max_frames = 0
    assert torch.equal(other, sep(mix[1:2]))
    sep.graph_max_frames = keep
    for _ in range(3):
        est, scores = st.run(mix, tgt)
        assert torch.equal(est, eager_run[0]) and torch.equal(scores, eager_run[1])
    assert st._run_graphs[(3, 9600)]["graph"] is not None
    # a larger call moves the workspace: graphs are dropped, results stay right
    gen = sep._ws_generation
    big = synthetic_mixture(2, 16000 * 200, seed=63).cuda()
    if sep.workspace_bytes(2, big.shape[1]) > sep._ws.numel():
        sep(big)
        assert sep._ws_generation == gen + 1 and not sep._graphs
    del big
    assert torch.equal(sep(mix[:1]), eager) and torch.equal(sep(mix[:1]), eager)
    est, scores = st.run(mix, tgt)
    assert torch.equal(est, eager_run[0]) and torch.equal(scores, eager_run[1])
    est, scores = st.run(mix, tgt)
    assert torch.equal(est, eager_run[0]) and torch.equal(scores, eager_run[1])


def test_streaming_step_batched_on_gpu(stage):
    """asr_audio_streaming for 6 concurrent streams through the real kernels: the batched step gives, stream by
    stream, what one-stream-at-a-time calls give (same decisions, same time ranges, same enrolment embeddings -
    batch-invariant bits), with one separator batch per step instead of one call per overlapped stream."""
    torch, st = stage
    from oracle import stream_toys as T
    from targetdiarization_b200 import streaming
    from targetdiarization_b200.synth import synthetic_mixture
    n_streams, n_steps = 6, 3
    g = np.random.default_rng(3)
    chunks = [[(synthetic_mixture(1, 9600, seed=100 + 10 * s + k)[0].numpy() * np.float32(0.2 + 0.1 * k))
               for k in range(n_streams)] for s in range(n_steps)]
    chunks[1][2] = np.zeros(9600, np.float32)                 # a silent chunk: gated out
    chunks[2][4] = chunks[2][4][:4800].copy()                 # under 0.4 s: ignored, clock not advanced
    overlap = [[False] * n_streams] + [[bool((s + k) % 2) for k in range(n_streams)] for s in range(1, n_steps)]
    kw = dict(asr=T.asr, vad=lambda a: T.vad(a), vad_inner=lambda a: T.vad(a, 0.0), similarity_threshold=0.4,
              separation_threshold=0.0)
    eng = streaming.StageEngine(st)

    def run(batched):
        states = [streaming.StreamState() for _ in range(n_streams)]
        out = []
        for s in range(n_steps):
            if batched:
                out.append(streaming.asr_audio_streaming_batch(eng, chunks[s], states, overlap[s], **kw))
            else:
                out.append([streaming.asr_audio_streaming_batch(eng, [chunks[s][k]], [states[k]], [overlap[s][k]],
                                                                **kw)[0] for k in range(n_streams)])
        return out, states
    a, sa = run(True)
    b, sb = run(False)
    assert a == b
    for x, y in zip(sa, sb):
        assert x.current_time == y.current_time and x.system_loudness_diff == y.system_loudness_diff
        assert np.array_equal(x.target_embedding, y.target_embedding)
    assert a[1][2] is None and a[2][4] is None
    assert sum(r is not None and r["type"] == "overlap" for row in a for r in row) >= 3
    assert sa[4].current_time == pytest.approx(1.2)
