"""CPU restatement (numpy / PyTorch) of the host logic around the separator: the chunk loop of
AudioProcessor.separate_speaker, the overlap-add of look2hear.utils.wav_chunk_inference, and the
target/non-target assignment of TargetASR.

TEST INFRASTRUCTURE ONLY (oracle): imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product path under targetdiarization_b200/.

Pinned by tests/golden/host_logic.npz, which oracle/make_golden.py produces by *executing the reference's own
functions* in the build container (AudioProcessor.separate_speaker extracted from the reference source with
`ast`, look2hear/utils/separator.py::wav_chunk_inference imported by path).  Paths below are relative to
/root/reference.
"""
import numpy as np
import torch

WINDOW = 160000  # AudioProcessor.py:896


def chunk_bounds(length, window=WINDOW):
    """[(start, end)] of the windows separate_speaker feeds to the separator (AudioProcessor.py:920-935, with
    vad_frame = [0, length]): n = L // W full windows; a remainder > W/2 becomes its own window, a smaller
    non-zero remainder extends the last window; L < W is one window."""
    n = length // window
    if n == 0:
        return [(0, length)]
    bounds = [(j * window, (j + 1) * window) for j in range(n)]
    rem = length % window
    if rem > 0:
        if rem > window / 2:
            bounds.append((bounds[-1][1], length))
        else:
            bounds[-1] = (bounds[-1][0], length)
    return bounds


def separate_speaker(audio, separater, loudness, window=WINDOW):
    """AudioProcessor.separate_speaker at 16 kHz without VAD (AudioProcessor.py:885-956).

    audio float32 [L]; separater: callable tensor [1,T] -> [1,2,T]; loudness: callable ndarray -> float
    (the reference uses pyloudnorm integrated LUFS rounded to 0.1, AudioProcessor.py:1123-1127).
    Returns (spk1, spk2) float32 [L], louder stream first (:949-952)."""
    spk1 = np.array([], dtype=np.float32)
    spk2 = np.array([], dtype=np.float32)
    for start, end in chunk_bounds(audio.shape[0], window):
        x = torch.from_numpy(audio[start:end].copy()).reshape(1, -1)
        with torch.no_grad():
            out = separater(x)
        if out.dim() == 3:
            out = out.squeeze(0)
        out = out.cpu().numpy()
        spk1 = np.concatenate([spk1, out[0]])
        spk2 = np.concatenate([spk2, out[1]])
    if loudness(spk1) < loudness(spk2):
        spk1, spk2 = spk2, spk1
    return spk1, spk2


def ola_plan(length, sr=16000, target_length=12.0, hop_length=4.0):
    """Segment plan of wav_chunk_inference (look2hear/utils/separator.py:84-101): returns
    (session, hop, pad, num_session, tr_ratio)."""
    session = int(sr * target_length)
    hop = int(sr * hop_length)
    pad = session - hop if session - hop > 0 else 0
    padded = length + 2 * pad
    num_session = (padded - session) // hop + 2
    return session, hop, pad, num_session, target_length / hop_length


def wav_chunk_inference(model, mixture, sr=16000, target_length=12.0, hop_length=4.0, batch_size=10, n_tracks=2):
    """look2hear/utils/separator.py:72-132 for ignore == 0 (session == target).

    mixture [1, nch, L]; model: [n, nch, session] -> [n, n_tracks, nch, session].  Returns [n_tracks, nch, L]."""
    L = mixture.shape[-1]
    session, hop, pad, num_session, tr_ratio = ola_plan(L, sr, target_length, hop_length)
    z = torch.zeros(mixture.shape[0], mixture.shape[1], pad, dtype=mixture.dtype)
    padded = torch.cat([z, mixture, z], -1) if pad > 0 else mixture
    acc = torch.zeros(padded.shape[0], n_tracks, padded.shape[1], padded.shape[2])
    segs, seglen = [], []
    for i in range(num_session):
        s = padded[:, :, i * hop:i * hop + session]
        n = s.shape[-1]
        if n < session:
            s = torch.cat([s, torch.zeros(s.shape[0], s.shape[1], session - n, dtype=s.dtype)], -1)
        segs.append(s)
        seglen.append(n)
    segs = torch.cat(segs, 0)
    for b0 in range(0, num_session, batch_size):
        with torch.no_grad():
            est = model(segs[b0:b0 + batch_size])
        for j in range(est.shape[0]):
            i = b0 + j
            acc[:, :, :, i * hop:i * hop + session] += est[j, :, :, :seglen[i]].unsqueeze(0)
    return (acc[:, :, :, pad:pad + L].contiguous() / tr_ratio).squeeze(0)


def cosine_similarity(a, b):
    """TargetASR.cosine_similarity (TargetASR.py:144-152): all-zero vector -> 1.0; clamp to [0,1]."""
    a = np.asarray(a)
    b = np.asarray(b)
    if np.all(a == 0.0) or np.all(b == 0.0):
        return 1.0
    s = np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))
    return float(max(0.0, min(s, 1.0)))


def pick_target(spk1_score, spk2_score, threshold=0.0):
    """multi_speakers_separate_asr / target_speaker_separate_asr (TargetASR.py:612-625, 541-553): None when
    both scores are below the threshold, else 1 iff spk1_score > spk2_score (strict), else 2."""
    if spk1_score < threshold and spk2_score < threshold:
        return None
    return 1 if spk1_score > spk2_score else 2


def pick_mix_audio(spk1_score, spk2_score, similarity_threshold=0.4):
    """mix_audio_processor (TargetASR.py:734-743), branch by branch: 0 = the unseparated input, 1 = spk1 (also on a
    tie: >=), 2 = spk2; the final else (every comparison false: NaN scores) returns the input.  Pinned by
    tests/golden/mix_rule.npz (the reference method run from source)."""
    if spk1_score < similarity_threshold and spk2_score < similarity_threshold:
        which = 0
    elif spk1_score >= spk2_score:
        which = 1
    elif spk2_score > spk1_score:
        which = 2
    else:
        which = 0
    return which


# ---------------------------------------------------------------------------------------------- loudness
def _k_weighting(rate):
    """BS.1770 K-weighting biquads as pyloudnorm (un-pinned dependency, requirements.txt:11 of the reference;
    absent from this image -> PARITY UNPINNED) builds them: high shelf (+4 dB, Q 1/sqrt2, 1500 Hz) and
    high pass (Q 0.5, 38 Hz), RBJ cookbook forms normalised by a0."""
    out = []
    for kind, G, Q, fc in (("high_shelf", 4.0, 1.0 / np.sqrt(2.0), 1500.0), ("high_pass", 0.0, 0.5, 38.0)):
        A = 10 ** (G / 40.0)
        w0 = 2.0 * np.pi * (fc / rate)
        alpha = np.sin(w0) / (2.0 * Q)
        c = np.cos(w0)
        if kind == "high_shelf":
            b = np.array([A * ((A + 1) + (A - 1) * c + 2 * np.sqrt(A) * alpha),
                          -2 * A * ((A - 1) + (A + 1) * c),
                          A * ((A + 1) + (A - 1) * c - 2 * np.sqrt(A) * alpha)])
            a = np.array([(A + 1) - (A - 1) * c + 2 * np.sqrt(A) * alpha,
                          2 * ((A - 1) - (A + 1) * c),
                          (A + 1) - (A - 1) * c - 2 * np.sqrt(A) * alpha])
        else:
            b = np.array([(1 + c) / 2, -(1 + c), (1 + c) / 2])
            a = np.array([1 + alpha, -2 * c, 1 - alpha])
        out.append((b / a[0], a / a[0]))
    return out


def integrated_loudness(audio, rate=16000, block=0.4):
    """Mono integrated loudness (LUFS), the algorithm of pyloudnorm.Meter.integrated_loudness: K-weighting,
    400 ms blocks with 75 % overlap, -70 LUFS absolute gate, -10 LU relative gate.  Raises ValueError on
    inputs shorter than one block, as pyloudnorm does."""
    from scipy.signal import lfilter
    x = np.asarray(audio, dtype=np.float64)
    if x.shape[0] < block * rate:
        raise ValueError("Audio must have length greater than the block size.")
    for b, a in _k_weighting(rate):
        x = lfilter(b, a, x)
    step = 0.25
    T = x.shape[0] / rate
    nblk = int(np.round((T - block) / (block * step)) + 1)
    z = np.zeros(nblk)
    for j in range(nblk):
        lo = int(block * (j * step) * rate)
        hi = int(block * (j * step + 1) * rate)
        z[j] = np.sum(np.square(x[lo:hi])) / (block * rate)
    with np.errstate(divide="ignore", invalid="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
        keep = l >= -70.0
        zg = np.mean(z[keep]) if keep.any() else np.nan
        gamma_r = -0.691 + 10.0 * np.log10(zg) - 10.0
        keep = (l > gamma_r) & (l > -70.0)
        zg = np.nan_to_num(np.mean(z[keep])) if keep.any() else 0.0
        return float(-0.691 + 10.0 * np.log10(zg))


def meter_loudness(audio, rate=16000):
    """AudioProcessor.meter_loudness (AudioProcessor.py:1123-1127): integrated LUFS rounded to 0.1."""
    return round(integrated_loudness(audio, rate), 1)


def get_target_embedding(audio_data_list, embed, cluster_labels, is_cluster=True, audio_input_type="separate",
                         output_embedding_list=True, sampling_rate=16000):
    """TargetASR.get_target_embedding after its preprocessing step (TargetASR.py:203-258), restated with the model
    calls as parameters: `embed(audio) -> [192]`, `cluster_labels(emb [n,192]) -> [n]` (-1 = noise).  Pinned by
    tests/golden/enrolment.npz (the reference method itself, run from source with the same stubs)."""
    audio_data_list = list(audio_data_list)
    if not audio_data_list:
        return np.zeros([192], dtype=np.float32)
    longest_audio = max(audio_data_list, key=lambda x: x.shape[0])
    normal_len_audios = [a for a in audio_data_list if a.shape[0] >= int(sampling_rate * 0.4)]
    merged_audio = audio_data_list[0] if len(audio_data_list) == 1 else np.concatenate(audio_data_list)
    if audio_input_type == "auto":
        if longest_audio.shape[0] >= 3.0 * sampling_rate:
            audio_input_type = "longest"
        elif len(normal_len_audios) <= 2:
            audio_input_type = "merge"
        else:
            audio_input_type = "separate"
    if audio_input_type == "merge":
        chosen = [merged_audio]
    elif audio_input_type == "longest":
        chosen = [longest_audio]
    else:
        chosen = normal_len_audios
    chosen = [a[:30 * sampling_rate] if a.shape[0] > 30 * sampling_rate else a for a in chosen]
    embedding_list = []
    for a in chosen:
        if a.shape[0] < 400:
            continue
        e = embed(a)
        if np.isnan(e).any():
            continue
        embedding_list.append(e)
    if is_cluster and len(embedding_list) > 2:
        labels = np.asarray(cluster_labels(np.stack(embedding_list)))
        valid = np.where(labels != -1)[0]
        if len(valid) > 0:
            embedding_list = [embedding_list[i] for i in valid]
    if output_embedding_list:
        return embedding_list
    if len(embedding_list) == 0:
        return np.zeros([192], dtype=np.float32)
    if len(embedding_list) == 1:
        return embedding_list[0]
    return np.mean(embedding_list, axis=0)
