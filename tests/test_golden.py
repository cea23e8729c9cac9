"""CPU tests: the oracle restatements and the product's host logic against the golden vectors that
oracle/make_golden.py produced by executing the reference (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import stage_port
from oracle.mossformer2_port import mossformer2_forward, snr_db
from targetdiarization_b200 import plan, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def host():
    return np.load(os.path.join(GOLDEN, "host_logic.npz"))


@pytest.fixture(scope="module")
def moss():
    return np.load(os.path.join(GOLDEN, "mossformer2_small.npz"))


def _golden_bounds(host):
    off = host["chunk_bounds_off"]
    flat = host["chunk_bounds_flat"]
    for i, L in enumerate(host["chunk_lengths"]):
        b = flat[off[i]:off[i + 1]].reshape(-1, 2)
        yield int(L), [(int(s), int(e)) for s, e in b]


def test_chunk_bounds_bit_exact(host):
    """Chunk boundaries of AudioProcessor.separate_speaker: oracle port and product planner vs the reference run."""
    n = 0
    for L, want in _golden_bounds(host):
        assert stage_port.chunk_bounds(L) == want, L
        assert plan.chunk_bounds(L) == want, L
        n += 1
    assert n == 18


def test_separate_speaker_port_matches_reference_run(host):
    from oracle.make_golden import rms_loudness, toy_separater
    seed, L = (int(v) for v in host["sepspk_audio_seed"])
    g = np.random.default_rng(seed)
    audio = (g.standard_normal(L) * 0.1).astype(np.float32)
    s1, s2 = stage_port.separate_speaker(audio, toy_separater, rms_loudness)
    assert s1.shape == (L,) and s2.shape == (L,)
    assert np.array_equal(s1[::997], host["sepspk_spk1_stride"])
    assert np.array_equal(s2[::997], host["sepspk_spk2_stride"])
    assert [tuple(b) for b in host["sepspk_bounds"].tolist()] == stage_port.chunk_bounds(L)


def test_wav_chunk_inference_port_matches_reference_run(host):
    mix = torch.from_numpy(host["ola_mix"])[None, None, :]

    def toy_model(x):
        t = torch.arange(x.shape[-1], dtype=torch.float32) / x.shape[-1]
        return torch.stack((x * (0.5 + t), x * x - 0.3 * t), dim=1)
    y = stage_port.wav_chunk_inference(toy_model, mix, sr=1000, target_length=12.0, hop_length=4.0, batch_size=10)
    assert np.array_equal(y[:, 0].numpy(), host["ola_out"])
    # the product's plan agrees on the segment count / geometry
    p = plan.ola_plan(mix.shape[-1], 1000, 12.0, 4.0)
    assert (p.session, p.hop, p.pad, p.ratio) == (12000, 4000, 8000, 3.0)
    assert p.num_session == stage_port.ola_plan(mix.shape[-1], 1000)[3]


def test_cosine_similarity_matches_reference_run(host):
    from targetdiarization_b200 import SeparationScoringStage
    a, tgt, want = host["cos_a"], host["cos_target"], host["cos_scores"]
    for i in range(len(a)):
        assert stage_port.cosine_similarity(a[i], tgt) == pytest.approx(want[i], abs=1e-7)
        assert SeparationScoringStage.cosine_similarity(a[i], tgt) == pytest.approx(want[i], abs=1e-7)
    assert want[2] == 1.0 and want[4] == 0.0 and want[5] == pytest.approx(1.0, abs=1e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_mossformer2_port_matches_reference_run(moss, tag):
    """oracle/mossformer2_port.py vs the reference module's output on the same weights and input (fp32 CPU both;
    fp32-vs-fp64 noise floor of the model is ~97 dB)."""
    seed, B, T = (int(v) for v in moss[f"{tag}_cfg"])
    sd = synth.random_state_dict(seed=seed, perturb=True)
    g = torch.Generator().manual_seed(1234 + seed)
    mix = torch.randn(B, T, generator=g) * 0.1
    taps = {}
    with torch.no_grad():
        y = mossformer2_forward(sd, mix, taps=taps)
    assert y.shape == (B, 2, T)
    assert snr_db(torch.from_numpy(moss[f"{tag}_out"]), y) >= 90.0
    for i in (0, 11, 23):
        for k in (f"flash{i}", f"layer{i}"):
            assert snr_db(torch.from_numpy(moss[f"{tag}_{k}"]), taps[k][:, ::37, ::7]) >= 90.0, k


def test_fbank_tables_match_kaldi():
    """The host-side tables the fbank kernel reads vs torchaudio's own (torchaudio/compliance/kaldi.py)."""
    import torchaudio.compliance.kaldi as kaldi
    from targetdiarization_b200 import fbank
    win = kaldi._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32)
    assert np.allclose(fbank.povey_window(), win.numpy(), atol=1e-7)
    mel, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    full, lo, hi = fbank.mel_banks()
    assert full.shape == (80, 257) and np.all(full[:, 256] == 0)
    assert np.array_equal(full[:, :256], mel.numpy())
    for i in range(80):
        assert np.all(full[i, :lo[i]] == 0) and np.all(full[i, hi[i] + 1:] == 0)
    assert fbank.num_frames(16037) == 98 and fbank.num_frames(399) == 0 and fbank.num_frames(400) == 1


def test_fbank_golden_is_kaldi():
    import torchaudio.compliance.kaldi as kaldi
    gd = np.load(os.path.join(GOLDEN, "fbank.npz"))
    seed, T = (int(v) for v in gd["cfg"])
    wav = synth.synthetic_mixture(1, T, seed=seed)
    f = kaldi.fbank(wav, num_mel_bins=80)
    f = f - f.mean(dim=0, keepdim=True)
    assert np.allclose(f.numpy(), gd["feat"], atol=1e-4)


def test_mossformer2_port_on_real_speech_excerpt():
    """Config C1 input (1.5 s of the reference's assets/chat_mix.wav) through the port vs the reference module."""
    gd = np.load(os.path.join(GOLDEN, "chat_mix_excerpt.npz"))
    sd = synth.random_state_dict(seed=0, perturb=True)
    x = torch.from_numpy(gd["pcm"].astype(np.float32) / 32768.0)[None]
    with torch.no_grad():
        y = mossformer2_forward(sd, x)
    assert snr_db(torch.from_numpy(gd["out"]), y) >= 90.0


# ---------------------------------------------------------------------------------------------- enrolment rules
def _toy_embedding(a):
    """The stub embedding oracle/make_golden.py::make_enrolment gave the reference method."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    e = np.array([a.size, a.sum(), a[:7].sum(), a[-5:].sum()] + [np.sin(a.size * (k + 1) * 1e-3) for k in range(188)])
    if a.size == 7777:
        e[3] = np.nan
    return e.astype(np.float32)


def test_enrolment_rules_match_reference_run():
    """plan.enrolment_select / enrolment_reduce against TargetASR.get_target_embedding executed from the reference
    source (tests/golden/enrolment.npz): selection mode, 0.4 s / 400-sample / 30 s rules, NaN skip, outlier drop,
    list and mean outputs - bit for bit."""
    from targetdiarization_b200 import plan
    z = np.load(os.path.join(GOLDEN, "enrolment.npz"))
    n_cases, n_labels = (int(v) for v in z["n_cases"])
    g = np.random.default_rng(int(z["seed"][0]))
    label_sets = [[0, 0, -1, 1, 1, -1], [-1, -1, -1, -1, -1, -1], [0, 0, 0, 0, 0, 0]]
    assert n_labels == len(label_sets)
    for ci in range(n_cases):
        lengths = [int(v) for v in z[f"c{ci}_lengths"]]
        mode = ["auto", "separate", "merge", "longest"][int(z[f"c{ci}_mode"][0])]
        pieces = [(g.standard_normal(n) * 0.1).astype(np.float32) for n in lengths]
        _, picks = plan.enrolment_select(lengths, 16000, mode)
        merged = np.concatenate(pieces) if len(pieces) > 1 else pieces[0]
        embs = [_toy_embedding((merged if src < 0 else pieces[src])[:n]) for src, n in picks]
        for li, labels in enumerate(label_sets):
            for cluster in (False, True):
                fn = lambda e, labels=labels: np.array(labels)[:len(e)]
                lst = plan.enrolment_reduce(embs, cluster, fn, True)
                one = plan.enrolment_reduce(embs, cluster, fn, False)
                key = f"c{ci}_l{li}_k{int(cluster)}"
                want = z[key + "_list"]
                assert len(lst) == want.shape[0], (key, len(lst), want.shape)
                if len(lst):
                    assert np.array_equal(np.stack(lst), want), key
                assert np.array_equal(np.asarray(one, dtype=np.float32), z[key + "_mean"]), key


def test_is_same_person_matches_reference_run():
    from oracle import stage_port
    from targetdiarization_b200 import plan
    z = np.load(os.path.join(GOLDEN, "enrolment.npz"))
    a, t = z["same_a"], z["same_t"]
    sim = stage_port.cosine_similarity(np.mean([a[i] for i in range(4)], axis=0), t)
    sim0 = stage_port.cosine_similarity(a[0], t)
    for row, thr in zip(z["same_res"], (0.2, 0.4, 0.9)):
        r = plan.is_same_person(sim, thr, verbose_result=True)
        assert float(r["is_same"]) == row[0] and r["score"] == row[1]
        assert float(plan.is_same_person(sim0, thr)) == row[2]


def test_oracle_enrolment_restatement_matches_reference_run():
    """oracle.stage_port.get_target_embedding against the same golden vectors (oracle pinned)."""
    from oracle import stage_port
    z = np.load(os.path.join(GOLDEN, "enrolment.npz"))
    n_cases, _ = (int(v) for v in z["n_cases"])
    g = np.random.default_rng(int(z["seed"][0]))
    label_sets = [[0, 0, -1, 1, 1, -1], [-1, -1, -1, -1, -1, -1], [0, 0, 0, 0, 0, 0]]
    for ci in range(n_cases):
        lengths = [int(v) for v in z[f"c{ci}_lengths"]]
        mode = ["auto", "separate", "merge", "longest"][int(z[f"c{ci}_mode"][0])]
        pieces = [(g.standard_normal(n) * 0.1).astype(np.float32) for n in lengths]
        for li, labels in enumerate(label_sets):
            for cluster in (False, True):
                fn = lambda e, labels=labels: np.array(labels)[:len(e)]
                lst = stage_port.get_target_embedding(pieces, _toy_embedding, fn, cluster, mode, True)
                one = stage_port.get_target_embedding(pieces, _toy_embedding, fn, cluster, mode, False)
                key = f"c{ci}_l{li}_k{int(cluster)}"
                want = z[key + "_list"]
                assert len(lst) == want.shape[0], key
                if len(lst):
                    assert np.array_equal(np.stack(lst), want), key
                assert np.array_equal(np.asarray(one, dtype=np.float32), z[key + "_mean"]), key


# ------------------------------------------------------------------------------- round 2: the BASELINE configurations
def test_mix_audio_processor_tie_rule_matches_reference_run():
    """TargetASR.mix_audio_processor (TargetASR.py:734-743) run from the reference source: which audio it returns for
    score pairs incl. exact ties (spk1 wins: >=), threshold edges and NaN scores, and the score it reports."""
    rows = np.load(os.path.join(GOLDEN, "mix_rule.npz"))["rows"]
    assert len(rows) == 28
    ties = 0
    for thr, a, b, which, score in rows:
        assert plan.pick_mix_audio(a, b, thr) == int(which), (thr, a, b)
        assert stage_port.pick_mix_audio(a, b, thr) == int(which)
        want = round(max(a, b), 3)
        assert (np.isnan(score) and np.isnan(want)) or score == want
        if a == b and not (a < thr):
            ties += 1
            assert int(which) == 1 and plan.pick_target(a, b, thr) == 2     # the two rules differ exactly on ties
    assert ties >= 4


def test_port_matches_reference_run_at_benchmark_shape():
    """oracle/mossformer2_port.py at T = 64 000 (item 0 of the C2 benchmark batch, 32 attention groups, 4 fixed-size
    linear-attention splits in the CUDA path) against the reference module's output on the same input."""
    gd = np.load(os.path.join(GOLDEN, "c2_item.npz"))
    seed, items, T, data_seed, _ = (int(v) for v in gd["cfg"])
    mix = synth.synthetic_mixture(items, T, seed=data_seed)[:1]
    with torch.no_grad():
        y = mossformer2_forward(synth.random_state_dict(seed=seed), mix)[0]
    assert snr_db(torch.from_numpy(gd["out_stride4"]), y[:, ::4]) >= 90.0


def test_port_and_embedder_oracle_on_the_whole_c1_file():
    """Config C1: the whole assets/chat_mix.wav through the port vs the reference run (>= 90 dB), and the oracle
    embedder's scores for the reference-separated streams are reproducible from the stored embeddings."""
    gd = np.load(os.path.join(GOLDEN, "c1_chat_mix.npz"))
    assert gd["mix_pcm"].shape == (138634,) and gd["target_pcm"].shape == (30768,)
    x = torch.from_numpy(gd["mix_pcm"].astype(np.float32) / 32768.0)[None]
    with torch.no_grad():
        y = mossformer2_forward(synth.random_state_dict(seed=0, perturb=True), x)[0]
    assert snr_db(torch.from_numpy(gd["out_stride4"]), y[:, ::4]) >= 90.0
    for k in range(2):
        assert stage_port.cosine_similarity(gd["emb"][k], gd["emb_target"]) == pytest.approx(gd["scores"][k], abs=1e-7)
    assert int(gd["pick"][0]) == (stage_port.pick_target(gd["scores"][0], gd["scores"][1], 0.0) or 0)
