"""GPU tests of the stage's host logic on top of the kernels: chunk loop (concat), overlap-add, scoring and
assignment, through the public Python surface that mirrors the reference methods."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(name, value):
    with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, "value": value}) + "\n")


@pytest.fixture(scope="module")
def stage():
    import torch
    from targetdiarization_b200 import SeparationScoringStage
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    return torch, SeparationScoringStage.random_init("cuda:0", seed=0)


def test_separate_speaker_concat_is_the_reference_chunk_loop(stage):
    """L = 400 001 -> windows [0,160000) [160000,320000) [320000,400001): the stitched streams equal, bit for bit,
    separator calls on exactly those windows; the louder stream comes first (AudioProcessor.py:949-952)."""
    torch, st = stage
    from oracle import stage_port
    from targetdiarization_b200.synth import synthetic_mixture
    L = 400001
    audio = synthetic_mixture(1, L, seed=21)[0].numpy()
    spk1, spk2 = st.separate_speaker(audio)
    assert spk1.shape == (L,) and spk2.shape == (L,) and spk1.dtype == np.float32
    bounds = stage_port.chunk_bounds(L)
    assert bounds == [(0, 160000), (160000, 320000), (320000, 400001)]
    parts = [st.separator(torch.from_numpy(audio[b:e]).cuda().unsqueeze(0))[0].cpu().numpy() for b, e in bounds]
    want = np.concatenate(parts, axis=1)
    l0, l1 = stage_port.meter_loudness(want[0]), stage_port.meter_loudness(want[1])
    if l0 < l1:
        want = want[::-1]
    assert np.array_equal(spk1, want[0]) and np.array_equal(spk2, want[1])
    assert stage_port.meter_loudness(spk1) >= stage_port.meter_loudness(spk2)


def test_separate_speaker_vs_cpu_oracle_low_ram_windows(stage):
    """1 s windows inside a VAD frame (low_gpu_ram path) against the CPU oracle end to end: >= 40 dB."""
    torch, st = stage
    from oracle import stage_port
    from oracle.mossformer2_port import mossformer2_forward, snr_db
    from targetdiarization_b200.synth import random_state_dict, synthetic_mixture
    sd = random_state_dict(seed=0)
    audio = synthetic_mixture(1, 43000, seed=22)[0].numpy()
    frames = [[2000, 42000]]
    s1, s2 = st.separate_speaker(audio, low_gpu_ram=True, vad_frames=frames, loudness=None)
    assert s1.shape == (42000,)   # the reference output ends at the last VAD frame
    assert not s1[:2000].any() and not s2[:2000].any()
    o1, o2 = stage_port.separate_speaker(audio[2000:42000], lambda x: mossformer2_forward(sd, x), lambda a: 0.0,
                                         window=16000)
    snr = snr_db(torch.from_numpy(np.stack((o1, o2))), torch.from_numpy(np.stack((s1[2000:], s2[2000:]))))
    _record("separate_speaker_low_ram_vs_oracle_snr_db", snr)
    assert snr >= 40.0


def test_wav_chunk_inference_is_bit_exact_overlap_add(stage):
    """The device gather/stitch kernels against the oracle's wav_chunk_inference driven by the same separator:
    identical segments in, identical sums out (ascending order, divide by 3)."""
    torch, st = stage
    from oracle import stage_port
    from targetdiarization_b200.synth import synthetic_mixture
    L = 30500
    mix = synthetic_mixture(1, L, seed=23)
    got = st.wav_chunk_inference(mix[None].cuda(), sr=1000).cpu()
    assert got.shape == (2, 1, L)

    def model(x):  # [n,1,session] -> [n,2,1,session]
        return st.separator(x[:, 0].cuda()).cpu().unsqueeze(2)
    want = stage_port.wav_chunk_inference(model, mix[None], sr=1000)
    assert torch.equal(got, want)


def test_separate_and_score_assignment(stage):
    """TargetASR.multi_speakers_separate_asr core: scores of both streams vs the target and the strict-> pick equal
    the oracle's on the same separated streams."""
    torch, st = stage
    from oracle import eres2netv2_port as E
    from oracle import stage_port
    from targetdiarization_b200.synth import random_eres2netv2_state_dict, synthetic_mixture
    esd = random_eres2netv2_state_dict(seed=0)
    audio = synthetic_mixture(1, 32000, seed=24)[0].numpy()
    target_wav = synthetic_mixture(1, 24000, seed=25)
    target = st.get_speaker_embedding(target_wav[0].numpy())
    assert target.shape == (192,) and target.dtype == np.float32
    r = st.separate_and_score(audio, target)
    ref_t = E.embed(esd, target_wav)[0].numpy()
    ref_e = E.embed(esd, torch.from_numpy(np.stack((r["spk1_audio"], r["spk2_audio"]))))
    s = [stage_port.cosine_similarity(ref_e[i].numpy(), ref_t) for i in range(2)]
    _record("score_abs_err_max", max(abs(s[0] - r["spk1_score"]), abs(s[1] - r["spk2_score"])))
    assert abs(s[0] - r["spk1_score"]) < 2e-3 and abs(s[1] - r["spk2_score"]) < 2e-3
    assert r["target"] == stage_port.pick_target(s[0], s[1], 0.0)
    assert st.separate_and_score(audio, target, threshold=1.5)["target"] is None


def test_run_batch_step(stage):
    torch, st = stage
    from targetdiarization_b200.synth import synthetic_mixture
    mix = synthetic_mixture(3, 16000, seed=26).cuda()
    tgt = st.embed(synthetic_mixture(1, 16000, seed=27).cuda())[0]
    est, scores = st.run(mix, tgt)
    assert est.shape == (3, 2, 16000) and scores.shape == (3, 2)
    again = st.score_segments(est.view(6, 16000), tgt).view(3, 2)
    assert torch.equal(scores, again)
    assert bool(((scores >= 0) & (scores <= 1)).all())
    from targetdiarization_b200 import Embedder, Separator
    assert st.launches_per_run(3, 16000) == Separator.KERNELS_PER_FORWARD + 2 + Embedder.KERNELS_PER_FORWARD + 1


def test_device_loudness_meter(stage):
    """BS.1770 integrated loudness on the device (fp64 K-weighting in independent warm-started segments + block
    means) against the oracle's pyloudnorm restatement: same value to 1e-6 dB, same 0.1-rounded value, on signals
    with silences (gating) and at several lengths incl. one just above the 400 ms minimum."""
    torch, st = stage
    import math
    from oracle import stage_port
    from targetdiarization_b200.pipeline import _block_bounds, _gate, _k_weighting
    from targetdiarization_b200.synth import synthetic_mixture
    for L, seed in ((6400 + 3, 1), (48000, 2), (160000 + 777, 3), (16000 * 70 + 5, 4)):
        x = synthetic_mixture(2, L, seed=seed)
        x[1] *= 0.05
        x[1, L // 3: L // 2] = 0.0     # silence: exercises the absolute gate
        got = st.meter_loudness_device(x.cuda())
        for i in range(2):
            want = stage_port.integrated_loudness(x[i].numpy())
            assert got[i] == round(want, 1), (L, i, got[i], want)
    with pytest.raises(ValueError):
        st.meter_loudness_device(torch.zeros(1, 3000).cuda())
    # known answer: full-scale 997 Hz sine = -3.0 LUFS (16 kHz RBJ filters: -3.06)
    t = torch.arange(48000) / 16000.0
    assert st.meter_loudness_device(torch.sin(2 * math.pi * 997.0 * t)[None].cuda())[0] == pytest.approx(-3.0, abs=0.15)


def test_full_size_c2_batch_properties(stage):
    """BASELINE config 2 at full size (64 mixtures x 4 s): size-independent properties instead of a CPU oracle run -
    every item of the batch equals, bit for bit, the same mixture separated on its own; scores are reproducible from
    the separated streams and lie in [0, 1]; nothing is NaN; the louder-first swap is a pure permutation."""
    torch, st = stage
    from targetdiarization_b200.synth import synthetic_mixture
    mix = synthetic_mixture(64, 64000, seed=31).cuda()
    tgt = st.embed(synthetic_mixture(1, 64000, seed=32).cuda())[0]
    est, scores = st.run(mix, tgt)
    assert est.shape == (64, 2, 64000) and scores.shape == (64, 2)
    assert bool(torch.isfinite(est).all()) and bool(torch.isfinite(scores).all())
    assert bool(((scores >= 0) & (scores <= 1)).all())
    for i in (0, 17, 63):
        single = st.separator(mix[i:i + 1])
        assert torch.equal(single[0], est[i]), f"item {i} depends on its batch"
    again = st.score_segments(est.view(128, 64000), tgt).view(64, 2)
    assert torch.equal(scores, again)
    # separate_speaker on one item returns the two streams of run(), louder first
    s1, s2 = st.separate_speaker(mix[5].cpu().numpy())
    a, b = est[5, 0].cpu().numpy(), est[5, 1].cpu().numpy()
    assert (np.array_equal(s1, a) and np.array_equal(s2, b)) or (np.array_equal(s1, b) and np.array_equal(s2, a))


def test_get_target_embedding_batched_equals_per_piece_loop(stage):
    """the enrolment path with ONE batched embedding call gives what the reference's per-piece loop
    gives (TargetASR.py:229-258) - same pieces selected, embeddings equal to per-piece calls, same mean."""
    torch, st = stage
    from targetdiarization_b200 import plan
    from targetdiarization_b200.synth import synthetic_mixture
    lengths = [9000, 6400, 399, 20000, 9000, 7000]
    pieces = [synthetic_mixture(1, n, seed=40 + i)[0].numpy() for i, n in enumerate(lengths)]
    labels = lambda e: np.array([0, 0, -1, 1, 1])[:len(e)]
    lst = st.get_target_embedding(pieces, is_preprocess=False, is_cluster=True, cluster_labels=labels)
    _, picks = plan.enrolment_select(lengths, 16000, "separate")
    assert [s for s, _ in picks] == [0, 1, 3, 4, 5]
    loop = [st.get_speaker_embedding(pieces[s][:n]) for s, n in picks]
    keep = [0, 1, 3, 4]
    assert len(lst) == len(keep)
    for got, k in zip(lst, keep):
        c = float(np.dot(got, loop[k]) / (np.linalg.norm(got) * np.linalg.norm(loop[k])))
        assert c > 0.99999, c
    one = st.get_target_embedding(pieces, is_preprocess=False, is_cluster=True, cluster_labels=labels,
                                  output_embedding_list=False)
    assert np.allclose(one, np.mean(lst, axis=0), atol=1e-6)
    # preprocessing hooks: VAD spans cut and concatenated, loudness callable applied, empty VAD -> piece dropped
    vad = lambda a: [] if a.shape[0] == 6400 else [[0.0, 0.25], [0.3, 0.5]]
    got = st.get_target_embedding(pieces[:2], is_preprocess=True, is_cluster=False, vad=vad,
                                  loudness_control=lambda a, sr: 0.5 * a, audio_input_type="merge")
    want = st.get_speaker_embedding(0.5 * np.concatenate([pieces[0][:4000], pieces[0][4800:8000]]))
    assert len(got) == 1 and np.allclose(got[0], want, atol=2e-3)
    # default clusterer (hdbscan / scikit-learn) runs on real embeddings
    assert 1 <= len(st.get_target_embedding(pieces, is_preprocess=False, is_cluster=True)) <= 5


def test_same_speaker_batch_equals_per_stream_rule4(stage):
    """rule 4 of the streaming gate for several streams in one batched call."""
    torch, st = stage
    from targetdiarization_b200 import plan
    from targetdiarization_b200.synth import synthetic_mixture
    prev = [synthetic_mixture(1, n, seed=60 + i)[0].numpy() for i, n in enumerate([19200, 9600, 28800])]
    cur = [synthetic_mixture(1, 9600, seed=70 + i)[0].numpy() for i in range(3)]
    cur[1] = prev[1].copy()           # identical audio -> similarity 1 -> same speaker
    got = st.same_speaker_batch(prev, cur, threshold=0.4, verbose_result=True)
    for i in range(3):
        sim = st.cosine_similarity(st.get_speaker_embedding(prev[i]), st.get_speaker_embedding(cur[i]))
        want = plan.is_same_person(sim, 0.4, verbose_result=True)
        assert got[i]["is_same"] == want["is_same"] and abs(got[i]["score"] - want["score"]) <= 0.002
    assert got[1]["is_same"] and got[1]["score"] >= 0.999


@pytest.mark.timeout(600)
def test_concurrent_python_threads_share_one_stage(stage):
    """SURVEY.md 8b threading: the reference shares ONE global model between its main thread and ThreadPoolExecutor
    workers without locking (main.py:42,338-367).  Four threads, each on its own CUDA stream, call the same stage -
    the graph-replayed small-call path (static buffers), the eager path (shared workspace) and the numpy call surface -
    and every result equals the serial one bit for bit."""
    torch, st = stage
    from concurrent.futures import ThreadPoolExecutor
    from targetdiarization_b200.synth import synthetic_mixture
    tgt = st.embed(synthetic_mixture(1, 64000, seed=32).cuda())[0]
    small = [synthetic_mixture(2, 9600, seed=200 + i).cuda() for i in range(4)]       # graphed: 2 x 1 199 frames
    big = [synthetic_mixture(24, 32000, seed=300 + i).cuda() for i in range(4)]       # eager: 24 x 4 096 frames
    host = [synthetic_mixture(1, 40000 + 1000 * i, seed=400 + i)[0].numpy() for i in range(4)]
    for _ in range(2):                                   # second pass captures the graphs of the small shape
        want_small = [tuple(t.clone() for t in st.run(m, tgt)) for m in small]
    want_big = [tuple(t.clone() for t in st.run(m, tgt)) for m in big]
    want_host = [st.separate_and_score(h, tgt, loudness=None) for h in host]
    torch.cuda.synchronize()

    def work(i):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        out = []
        with torch.cuda.stream(s):
            for rep in range(3):
                a = st.run(small[i], tgt)
                b = st.run(big[i], tgt)
                c = st.separate_and_score(host[i], tgt, loudness=None)
                out.append((a, b, c))
            s.synchronize()
        return out

    with ThreadPoolExecutor(max_workers=4) as pool:
        results = list(pool.map(work, range(4)))
    torch.cuda.synchronize()
    for i, reps in enumerate(results):
        for a, b, c in reps:
            assert torch.equal(a[0], want_small[i][0]) and torch.equal(a[1], want_small[i][1]), f"thread {i}: graphed path"
            assert torch.equal(b[0], want_big[i][0]) and torch.equal(b[1], want_big[i][1]), f"thread {i}: eager path"
            assert np.array_equal(c["spk1_audio"], want_host[i]["spk1_audio"]), f"thread {i}: numpy surface"
            assert c["spk1_score"] == want_host[i]["spk1_score"] and c["target"] == want_host[i]["target"]
