"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE in the build container (where /root/reference exists).

TEST INFRASTRUCTURE ONLY (oracle).  Run:  python -m oracle.make_golden
The vectors are small (a few hundred KB in total) and are committed together with this script, so that the
oracle restatements (oracle/mossformer2_port.py, oracle/stage_port.py) and the CUDA path can be checked against
outputs of the reference itself on the GPU box, where the reference tree does not exist.

  mossformer2_small.npz   reference MossFormer2 module (look2hear/models/mossformer2.py) loaded with the weights of
                          targetdiarization_b200.synth.random_state_dict(seed, perturb=True): full forward on two
                          ragged inputs + layer taps (strided samples, to stay small)
  host_logic.npz          AudioProcessor.separate_speaker (extracted from AudioProcessor.py with `ast`, run with a
                          recording stand-in for self.separater) -> chunk boundaries for a list of lengths;
                          look2hear.utils.wav_chunk_inference (imported by path) with a toy model -> stitched output;
                          TargetASR.cosine_similarity (extracted with `ast`) on fixed vectors
  enrolment.npz           TargetASR.get_target_embedding (is_preprocess=False) and TargetASR.is_same_person, extracted
                          with `ast` and run with stub models (deterministic toy embedding, fixed cluster labels):
                          selection mode, 0.4 s / 400-sample / 30 s rules, NaN skip, outlier drop, list / mean outputs
  chat_mix_excerpt.npz    1.5 s of assets/chat_mix.wav (the reference's demo input, config C1) through the reference
                          MossFormer2 module with the seed-0 synthetic weights
  fbank.npz               torchaudio.compliance.kaldi.fbank (the function the modelscope pipeline calls) on a fixed
                          1 s signal, mean-normalised
  c1_chat_mix.npz         BASELINE config 1 end to end: the WHOLE assets/chat_mix.wav (8.665 s, S = 17 328 frames)
                          through the reference MossFormer2 module (seed-0 perturbed weights; every 4th output sample
                          kept), assets/female_a.wav as the target, and the embeddings / cosine scores / pick the
                          oracle embedder gives for the reference-separated streams
  c2_item.npz             BASELINE config 2: item 0 of the benchmark batch (synthetic_mixture(64, 64000, seed=1234),
                          the bench's seed-0 weights) through the reference module at T = 64 000 (every 4th sample)
  streaming.npz           TargetDiarizationStream.asr_audio_streaming + TargetASR.multi_speakers_separate_asr /
                          single_speaker_asr / is_same_person (all extracted with `ast`, toy models of
                          oracle/stream_toys.py): 7 steps x 4 concurrent streams, results and per-stream state
  mix_rule.npz            TargetASR.mix_audio_processor (extracted with `ast`, stub models) on score pairs incl. ties
                          and NaN: which audio it returns and the score it reports (the >= tie rule, TargetASR.py:734-743)
  apollo_small.npz        SURVEY.md 8f-4: the reference Apollo module (look2hear/models/apollo.py, the configuration of
                          AudioProcessor.py:279) loaded with synth.random_apollo_state_dict(0) on a [1, 2, 22 173]
                          full-band signal (T = 51 frames) and on a [2, 1, 4 500] one: the restored waveforms
  mdx_stft.npz            the reference's ConvTDFNet (AudioProcessor.py:65-120, compiled from its source lines):
                          stft of two stereo chunks at n_fft 6144 / hop 1024 / dim_f 3072 / dim_t 8 (every 5th bin kept)
                          and istft of a fixed random spectrogram
"""
import ast
import io
import os
import sys
import textwrap
import types
import typing

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from targetdiarization_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CHUNK_LENGTHS = [1, 15999, 79999, 80000, 80001, 159999, 160000, 160001, 239999, 240000, 240001, 320000, 400000,
                 400001, 480000, 559999, 1000000, 1680001]


def extract_method(path, cls, name):
    """Source of method `name` of class `cls` in the reference file `path`, dedented."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == name:
                    return textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
    raise KeyError(f"{cls}.{name} not found in {path}")


def reference_separate_speaker():
    """AudioProcessor.separate_speaker compiled from the reference source; returns a function
    (audio, separater, loudness) -> (spk1, spk2, [(start, end)...])."""
    code = extract_method(os.path.join(ref_loader.REF_ROOT, "AudioProcessor.py"), "AudioProcessor", "separate_speaker")
    ns = {"np": np, "torch": torch}
    exec(compile(code, "AudioProcessor.separate_speaker", "exec"), ns)
    fn = ns["separate_speaker"]

    def run(audio, separater, loudness):
        calls = []

        def rec(x):
            calls.append(int(x.shape[-1]))
            return separater(x)
        fake = types.SimpleNamespace(
            is_separate_audio=True, verbose_log=False, device="cpu", separater=rec,
            ndarray_to_torchaudio=lambda a: torch.tensor(a.reshape(1, -1)),   # AudioProcessor.py:1051-1056
            meter_loudness=lambda audio_data, sampling_rate: loudness(audio_data))
        s1, s2 = fn(fake, audio, 16000, False)
        bounds, pos = [], 0
        for n in calls:
            bounds.append((pos, pos + n))
            pos += n
        return s1, s2, bounds
    return run


def toy_separater(x):
    """[1,T] -> [1,2,T]: cheap, chunk-position dependent (so boundaries show in the output)."""
    t = torch.arange(x.shape[-1], dtype=torch.float32) / x.shape[-1]
    return torch.stack((x * (0.5 + t), -0.25 * x + 0.1 * t), dim=1)


def rms_loudness(a):
    return round(float(10 * np.log10(np.mean(np.square(a.astype(np.float64))) + 1e-12)), 1)


def make_host_logic():
    run = reference_separate_speaker()
    out = {}
    bounds_flat, bounds_off = [], [0]
    for L in CHUNK_LENGTHS:
        audio = np.zeros(L, dtype=np.float32)
        _, _, b = run(audio, lambda x: torch.zeros(1, 2, x.shape[-1]), lambda a: 0.0)
        bounds_flat += [v for se in b for v in se]
        bounds_off.append(len(bounds_flat))
    out["chunk_lengths"] = np.array(CHUNK_LENGTHS, dtype=np.int64)
    out["chunk_bounds_flat"] = np.array(bounds_flat, dtype=np.int64)
    out["chunk_bounds_off"] = np.array(bounds_off, dtype=np.int64)
    # full run with a toy separater on 2.6 windows; stream order decided by the stub loudness
    g = np.random.default_rng(7)
    audio = (g.standard_normal(416000) * 0.1).astype(np.float32)
    s1, s2, b = run(audio, toy_separater, rms_loudness)
    out["sepspk_audio_seed"] = np.array([7, 416000], dtype=np.int64)
    out["sepspk_spk1_stride"] = s1[::997].copy()
    out["sepspk_spk2_stride"] = s2[::997].copy()
    out["sepspk_bounds"] = np.array(b, dtype=np.int64)
    # wav_chunk_inference with a toy model at sr = 1000 (session 12000, hop 4000)
    wci = ref_loader.load_reference_wav_chunk_inference()
    mix = torch.from_numpy((g.standard_normal(30500) * 0.1).astype(np.float32))[None, None, :]

    def toy_model(x):  # [n,1,session] -> [n,2,1,session]
        t = torch.arange(x.shape[-1], dtype=torch.float32) / x.shape[-1]
        return torch.stack((x * (0.5 + t), x * x - 0.3 * t), dim=1)
    y = wci(toy_model, mix, sr=1000, target_length=12.0, hop_length=4.0, batch_size=10, n_tracks=2)
    out["ola_mix"] = mix[0, 0].numpy()
    out["ola_out"] = y[:, 0].numpy()
    # cosine_similarity from TargetASR
    code = extract_method(os.path.join(ref_loader.REF_ROOT, "TargetASR.py"), "TargetASR", "cosine_similarity")
    ns = {"np": np}
    exec(compile(code, "TargetASR.cosine_similarity", "exec"), ns)
    a = g.standard_normal((6, 192)).astype(np.float32)
    tgt = g.standard_normal(192).astype(np.float32)
    a[2] = 0.0
    a[4] = -tgt
    a[5] = 3 * tgt
    out["cos_a"] = a
    out["cos_target"] = tgt
    out["cos_scores"] = np.array([ns["cosine_similarity"](None, a[i], tgt) for i in range(6)], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "host_logic.npz"), **out)
    print("host_logic.npz:", {k: v.shape for k, v in out.items()})


def make_enrolment():
    """TargetASR.get_target_embedding (is_preprocess=False path) and TargetASR.is_same_person, executed from the
    reference source with stubs: the embedding of a piece is a deterministic function of its samples, the
    clusterer labels a fixed pattern, so only the reference's own selection / truncation / reduction rules act."""
    path = os.path.join(ref_loader.REF_ROOT, "TargetASR.py")
    ns = {"np": np, "Union": typing.Union, "Literal": typing.Literal, "io": io}
    labels_for = {}

    class _Clusterer:
        def __init__(self, **kw):
            assert kw == {"min_cluster_size": 2, "metric": "euclidean"}, kw

        def fit_predict(self, emb):
            return labels_for["labels"][:len(emb)]
    ns["hdbscan"] = types.SimpleNamespace(HDBSCAN=_Clusterer)
    for name in ("get_target_embedding", "is_same_person", "cosine_similarity"):
        exec(compile(extract_method(path, "TargetASR", name), "TargetASR." + name, "exec"), ns)

    def toy_embedding(wav_file, embedding_model="eres2netv2_large"):
        a = np.asarray(wav_file, dtype=np.float64).reshape(-1)
        e = np.array([a.size, a.sum(), a[:7].sum(), a[-5:].sum()] + [np.sin(a.size * (k + 1) * 1e-3) for k in range(188)])
        if a.size == 7777:          # one piece yields a NaN embedding (must be skipped)
            e[3] = np.nan
        return e.astype(np.float32)
    fake = types.SimpleNamespace(
        verbose_log=False, get_speaker_embedding=toy_embedding,
        ap=types.SimpleNamespace(combine_audio_chunks=lambda audio_data_list: audio_data_list[0]
                                 if len(audio_data_list) == 1 else np.concatenate(audio_data_list)))
    fake.cosine_similarity = lambda embedding_a, embedding_b: ns["cosine_similarity"](fake, embedding_a, embedding_b)
    g = np.random.default_rng(11)
    cases = [
        ([3000, 9000, 20000, 399, 6400, 7777], "separate"),
        ([3000, 9000, 20000, 399, 6400, 7777], "auto"),
        ([3000, 5000], "auto"),
        ([50000, 3000, 50000, 100], "auto"),
        ([500000, 3000], "longest"),
        ([300000, 300000, 3000], "merge"),
        ([100, 200, 399], "separate"),
        ([100, 200], "merge"),
        ([6400, 6401, 6402, 6403, 6404], "separate"),
    ]
    label_sets = [[0, 0, -1, 1, 1, -1], [-1, -1, -1, -1, -1, -1], [0, 0, 0, 0, 0, 0]]
    out = {"n_cases": np.array([len(cases), len(label_sets)], dtype=np.int64), "seed": np.array([11], dtype=np.int64)}
    for ci, (lengths, mode) in enumerate(cases):
        pieces = [(g.standard_normal(n) * 0.1).astype(np.float32) for n in lengths]
        out[f"c{ci}_lengths"] = np.array(lengths, dtype=np.int64)
        out[f"c{ci}_mode"] = np.array([["auto", "separate", "merge", "longest"].index(mode)], dtype=np.int64)
        for li, labels in enumerate(label_sets):
            labels_for["labels"] = np.array(labels)
            for cluster in (False, True):
                # the method takes file paths in a list: the stub's read_audio / audio_resample hand back the prepared
                # arrays, so the pieces reach the selection rules exactly as after the reference's own reading step
                fake.ap.read_audio = lambda file_path: (pieces[int(file_path)], 16000)
                fake.ap.audio_resample = lambda audio_data, orig_sr, target_sr: (audio_data, None)
                lst = ns["get_target_embedding"](fake, [str(i) for i in range(len(pieces))], False, cluster,
                                                 "eres2netv2_large", mode, True)
                one = ns["get_target_embedding"](fake, [str(i) for i in range(len(pieces))], False, cluster,
                                                 "eres2netv2_large", mode, False)
                key = f"c{ci}_l{li}_k{int(cluster)}"
                out[key + "_list"] = np.stack(lst) if len(lst) else np.zeros((0, 192), np.float32)
                out[key + "_mean"] = np.asarray(one, dtype=np.float32)
    # is_same_person
    a = g.standard_normal((4, 192)).astype(np.float32)
    t = (a.mean(0) + 0.8 * g.standard_normal(192)).astype(np.float32)
    res = []
    for thr in (0.2, 0.4, 0.9):
        r = ns["is_same_person"](fake, [a[i] for i in range(4)], t, thr, True)
        res.append([float(r["is_same"]), r["score"], float(ns["is_same_person"](fake, a[0], t, thr, False))])
    out["same_a"] = a
    out["same_t"] = t
    out["same_res"] = np.array(res, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "enrolment.npz"), **out)
    print("enrolment.npz:", len(out), "arrays")


def make_mossformer2():
    pkg = ref_loader.load_reference_modules()
    out = {}
    for tag, seed, B, T in (("a", 0, 1, 4000), ("b", 1, 2, 2413)):
        sd = synth.random_state_dict(seed=seed, perturb=True)
        model = pkg.mossformer2.MossFormer2().eval()
        missing, unexpected = model.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(1234 + seed)
        mix = torch.randn(B, T, generator=g) * 0.1
        taps = {}
        hooks = []
        blk = model.mask_net.mdl.intra_mdl.mossformerM
        for i in (0, 11, 23):
            hooks.append(blk.layers[i].register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"flash{i}", o)))
            hooks.append(blk.fsmn[i].register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"layer{i}", o)))
        with torch.no_grad():
            y = model(mix)
        for h in hooks:
            h.remove()
        out[f"{tag}_cfg"] = np.array([seed, B, T], dtype=np.int64)
        out[f"{tag}_out"] = y.numpy()
        for k, v in taps.items():
            out[f"{tag}_{k}"] = v[:, ::37, ::7].contiguous().numpy()   # strided sample of [B,S,512]
    np.savez_compressed(os.path.join(GOLDEN, "mossformer2_small.npz"), **out)
    print("mossformer2_small.npz:", {k: v.shape for k, v in out.items()})


def make_chat_mix_excerpt():
    """1.5 s of the reference's demo mixture assets/chat_mix.wav (config C1 input; int16 PCM, samples 32000..55999)
    through the reference MossFormer2 module with the seed-0 synthetic weights: real speech statistics."""
    from scipy.io import wavfile
    sr, wav = wavfile.read(os.path.join(ref_loader.REF_ROOT, "assets", "chat_mix.wav"))
    assert sr == 16000 and wav.dtype == np.int16
    ex = wav[32000:32000 + 24000].copy()
    pkg = ref_loader.load_reference_modules()
    model = pkg.mossformer2.MossFormer2().eval()
    model.load_state_dict(synth.random_state_dict(seed=0, perturb=True), strict=True)
    x = torch.from_numpy(ex.astype(np.float32) / 32768.0)[None]   # AudioProcessor.int16_to_float32 (:1043-1048)
    with torch.no_grad():
        y = model(x)
    np.savez_compressed(os.path.join(GOLDEN, "chat_mix_excerpt.npz"), pcm=ex, offset=np.array([32000]), out=y.numpy())
    print("chat_mix_excerpt.npz:", tuple(y.shape))


def make_fbank():
    import torchaudio.compliance.kaldi as kaldi
    wav = synth.synthetic_mixture(1, 16037, seed=5)
    f = kaldi.fbank(wav, num_mel_bins=80)
    f = f - f.mean(dim=0, keepdim=True)
    np.savez_compressed(os.path.join(GOLDEN, "fbank.npz"), cfg=np.array([5, 16037], dtype=np.int64), feat=f.numpy())
    print("fbank.npz:", tuple(f.shape))


def make_c1_chat_mix():
    """Config C1 (target_diarization_test.py:26-40): the reference's demo mixture and target sample."""
    from scipy.io import wavfile
    from oracle import eres2netv2_port as E
    from oracle import stage_port
    sr, wav = wavfile.read(os.path.join(ref_loader.REF_ROOT, "assets", "chat_mix.wav"))
    sr2, tgt = wavfile.read(os.path.join(ref_loader.REF_ROOT, "assets", "female_a.wav"))
    assert sr == 16000 and sr2 == 16000 and wav.dtype == np.int16 and tgt.dtype == np.int16
    pkg = ref_loader.load_reference_modules()
    model = pkg.mossformer2.MossFormer2().eval()
    model.load_state_dict(synth.random_state_dict(seed=0, perturb=True), strict=True)
    x = torch.from_numpy(wav.astype(np.float32) / 32768.0)[None]   # AudioProcessor.int16_to_float32 (:1043-1048)
    with torch.no_grad():
        y = model(x)[0]                                             # [2, 138634]
    esd = E.random_state_dict(seed=0)
    t = torch.from_numpy(tgt.astype(np.float32) / 32768.0)[None]
    emb = E.embed(esd, torch.cat((y, torch.nn.functional.pad(t, (0, y.shape[1] - t.shape[1])))))[:2]
    emb_t = E.embed(esd, t)[0]
    scores = [stage_port.cosine_similarity(emb[k].numpy(), emb_t.numpy()) for k in range(2)]
    np.savez_compressed(os.path.join(GOLDEN, "c1_chat_mix.npz"), mix_pcm=wav, target_pcm=tgt, out_stride4=y[:, ::4].numpy(),
                        emb=emb.numpy(), emb_target=emb_t.numpy(), scores=np.array(scores, dtype=np.float64),
                        pick=np.array([stage_port.pick_target(scores[0], scores[1], 0.0) or 0], dtype=np.int64))
    print("c1_chat_mix.npz:", tuple(y.shape), "scores", scores)


def make_c2_item():
    pkg = ref_loader.load_reference_modules()
    model = pkg.mossformer2.MossFormer2().eval()
    model.load_state_dict(synth.random_state_dict(seed=0), strict=True)
    mix = synth.synthetic_mixture(64, 64000, seed=1234)[:1]
    with torch.no_grad():
        y = model(mix)[0]
    np.savez_compressed(os.path.join(GOLDEN, "c2_item.npz"), cfg=np.array([0, 64, 64000, 1234, 0], dtype=np.int64),
                        out_stride4=y[:, ::4].numpy())
    print("c2_item.npz:", tuple(y.shape))


def make_mix_rule():
    """TargetASR.mix_audio_processor's choice between the unseparated input, spk1 and spk2."""
    path = os.path.join(ref_loader.REF_ROOT, "TargetASR.py")
    ns = {"np": np, "Union": typing.Union, "io": io}
    for name in ("mix_audio_processor",):
        exec(compile(extract_method(path, "TargetASR", name), "TargetASR." + name, "exec"), ns)
    audio = np.full(16000, 0.25, dtype=np.float32)
    s1, s2 = np.full(16000, 1.0, dtype=np.float32), np.full(16000, 2.0, dtype=np.float32)
    cases = [(0.5, 0.5), (0.7, 0.2), (0.2, 0.7), (0.3, 0.3), (0.39999, 0.4), (0.4, 0.39999), (0.1, 0.2),
             (float("nan"), 0.9), (0.9, float("nan")), (float("nan"), float("nan")), (1.0, 1.0), (0.0, 0.0),
             (0.4, 0.4), (0.45, 0.450001)]
    rows = []
    for thr in (0.4, 0.0):
        for a, b in cases:
            scores = iter([a, b])
            fake = types.SimpleNamespace(
                input_audio_preprocess=lambda audio: (audio, 16000),
                mdx_weights_file=None,
                ap=types.SimpleNamespace(meter_loudness=lambda audio_data, sampling_rate: -20.0,
                                         denoise_vocal=lambda audio_data, sampling_rate: audio_data,
                                         audio_loudness_control=lambda audio_data, sampling_rate: audio_data,
                                         separate_speaker=lambda audio_data: (s1, s2)),
                asrp=types.SimpleNamespace(is_diarization=True,
                                           speaker_diarization=lambda wav_file, sampling_rate: [1, 2]),
                get_speaker_embedding=lambda wav_file, embedding_model="eres2netv2_large": wav_file[:1],
                cosine_similarity=lambda embedding_a, embedding_b: next(scores))
            r = ns["mix_audio_processor"](fake, audio, np.zeros(192, np.float32), thr, -40.0)
            which = {0.25: 0, 1.0: 1, 2.0: 2}[float(r["audio"][0])]
            rows.append([thr, a, b, which, r["score"]])
    np.savez_compressed(os.path.join(GOLDEN, "mix_rule.npz"), rows=np.array(rows, dtype=np.float64))
    print("mix_rule.npz:", len(rows), "cases")


def make_streaming():
    """TargetDiarizationStream.asr_audio_streaming and the TargetASR methods it calls (multi_speakers_separate_asr,
    single_speaker_asr, is_same_person, cosine_similarity), all extracted from the reference source and run on the
    scenario of oracle/stream_toys.py with the toy models: per call the returned dict (or None) and the stream state."""
    import re
    from oracle import stream_toys as T
    tasr_path = os.path.join(ref_loader.REF_ROOT, "TargetASR.py")
    stream_path = os.path.join(ref_loader.REF_ROOT, "TargetDiarizationStream.py")
    ns = {"np": np, "Union": typing.Union, "Literal": typing.Literal, "io": io, "re": re}
    for name in ("multi_speakers_separate_asr", "single_speaker_asr", "is_same_person", "cosine_similarity"):
        exec(compile(extract_method(tasr_path, "TargetASR", name), "TargetASR." + name, "exec"), ns)
    exec(compile(extract_method(stream_path, "TargetDiarizationStream", "asr_audio_streaming"),
                 "TargetDiarizationStream.asr_audio_streaming", "exec"), ns)

    def make_stream():
        tasr = types.SimpleNamespace(
            mdx_weights_file=None, restorer_weights_folder=None, silero_vad=None,
            input_audio_preprocess=lambda audio: (audio, 16000),
            get_speaker_embedding=lambda wav_file, embedding_model="eres2netv2_large": T.embedding(wav_file),
            ap=types.SimpleNamespace(separate_speaker=lambda audio_data: T.separate_speaker(audio_data)),
            asrp=types.SimpleNamespace(
                vad_detection=lambda wav_file, min_silence_sec=None: T.vad(wav_file, min_silence_sec),
                asr_detection=lambda wav_file, asr_engine, prompt, output_text_only, no_punc: T.asr(wav_file, prompt)))
        for name in ("multi_speakers_separate_asr", "single_speaker_asr", "is_same_person", "cosine_similarity"):
            setattr(tasr, name, types.MethodType(ns[name], tasr))
        return types.SimpleNamespace(
            current_time=0.0, target_embedding=None, prev_asr_text="", system_loudness_diff=0.0,
            use_asr_prompt=True, similarity_threshold=0.4, loudness_diff_threshold=12.0, tasr=tasr,
            ap=types.SimpleNamespace(meter_loudness=lambda audio_data, sampling_rate: T.meter_loudness(audio_data)),
            audio_preprocess=lambda audio_data, sampling_rate, stream_mode, output_audio_only: T.audio_preprocess(audio_data))

    chunks, overlap = T.scenario()
    n_steps, n_streams = len(chunks), len(chunks[0])
    streams = [make_stream() for _ in range(n_streams)]
    rec = {"shape": np.array([n_steps, n_streams], dtype=np.int64)}
    texts = []
    for s in range(n_steps):
        for k in range(n_streams):
            r = ns["asr_audio_streaming"](streams[k], chunks[s][k], overlap[s][k])
            if r is not None:                       # process_single_chunk (:184-186)
                streams[k].prev_asr_text = r["text"]
            st = streams[k]
            te = np.zeros(192, np.float32) if st.target_embedding is None else st.target_embedding
            rec[f"r{s}_{k}"] = np.array([0.0] if r is None else
                                        [1.0, float(r["speaker"]), r["timerange"][0], r["timerange"][1],
                                         float(r["type"] == "overlap")], dtype=np.float64)
            rec[f"s{s}_{k}"] = np.array([st.current_time, st.system_loudness_diff, float(te.sum()),
                                         float(np.abs(te).sum())], dtype=np.float64)
            texts.append("" if r is None else r["text"])
    rec["texts"] = np.array(texts)
    np.savez_compressed(os.path.join(GOLDEN, "streaming.npz"), **rec)
    n_res = sum(1 for t in texts if t)
    print("streaming.npz:", n_steps, "steps x", n_streams, "streams,", n_res, "results")


def make_apollo():
    sd = synth.random_apollo_state_dict(0)
    m = ref_loader.build_reference_apollo(sd)
    xa = synth.synthetic_fullband(2, 22050 + 123, seed=4321).reshape(1, 2, -1)
    xb = synth.synthetic_fullband(2, 4500, seed=77).reshape(2, 1, -1)
    with torch.no_grad():
        ya, yb = m(xa), m(xb)
    np.savez_compressed(os.path.join(GOLDEN, "apollo_small.npz"), out_a=ya.numpy(), out_b=yb.numpy())


def make_mdx():
    cls = ref_loader.load_reference_conv_tdf_net()
    net = cls(target_name="vocals", L=11, dim_f=3072, dim_t=3, n_fft=6144, hop=1024, device="cpu")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 2, net.chunk_size, generator=g) * 0.1
    spec_in = torch.randn(2, 4, 3072, 8, generator=g)
    with torch.no_grad():
        spec = net.stft(x)
        wav = net.istft(spec_in)
    np.savez_compressed(os.path.join(GOLDEN, "mdx_stft.npz"), x=x.numpy(), spec_bins5=spec[:, :, ::5].numpy(),
                        spec_in_seed=np.int64(5), wav=wav.numpy(), chunk_size=np.int64(net.chunk_size))


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("the reference tree is not present; golden vectors can only be generated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    make_host_logic()
    make_enrolment()
    make_fbank()
    make_mossformer2()
    make_chat_mix_excerpt()
    make_c1_chat_mix()
    make_c2_item()
    make_mix_rule()
    make_streaming()
    make_apollo()
    make_mdx()
