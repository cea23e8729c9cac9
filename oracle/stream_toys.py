"""Deterministic stand-ins for the models around TargetDiarizationStream.asr_audio_streaming, shared by the golden
generator (oracle/make_golden.py::make_streaming, where they are plugged into the REFERENCE method run from its
source) and by tests/test_streaming.py (where the same functions drive the product's batched orchestration).

TEST INFRASTRUCTURE ONLY (oracle)."""
import numpy as np


def meter_loudness(audio_data, sampling_rate=16000):
    """Stand-in for AudioProcessor.meter_loudness: RMS level in dB, rounded to 0.1 like the reference's value."""
    a = np.asarray(audio_data, dtype=np.float64)
    return round(float(10 * np.log10(np.mean(a * a) + 1e-12)), 1)


def audio_preprocess(audio_data):
    return (np.asarray(audio_data, dtype=np.float32) * np.float32(0.9)).astype(np.float32)


def embedding(wav):
    """192-d 'speaker embedding': normalised autocorrelation lags + a few moments (varies with the content)."""
    a = np.asarray(wav, dtype=np.float64).reshape(-1)
    n = a.size
    lags = np.array([np.dot(a[:n - k], a[k:]) / max(n - k, 1) for k in range(1, 189)])
    e = np.concatenate(([a.mean(), np.abs(a).mean(), a.std(), float(n) * 1e-6], lags / (a.var() + 1e-9)))
    return e.astype(np.float32)


def vad(audio, min_silence_sec=None):
    """Speech spans: the whole clip unless it is (nearly) silent; the inner form (min_silence_sec=0.0) trims 10 ms."""
    a = np.asarray(audio, dtype=np.float64)
    dur = round(a.size / 16000, 3)
    if a.size == 0 or np.sqrt(np.mean(a * a)) < 1e-3:
        return []
    return [[0.01, dur]] if min_silence_sec is not None else [[0.0, dur]]


def asr(audio, prompt=""):
    """'Transcript': a word whose length depends on the clip's energy, with punctuation; empty for very weak clips."""
    a = np.asarray(audio, dtype=np.float64)
    rms = float(np.sqrt(np.mean(a * a))) if a.size else 0.0
    if rms < 3e-3:
        return "  "
    k = 1 + int(rms * 400) % 7
    return ("Ab" * k) + ("!" if k % 2 else ", ok.") + (f" [{len(prompt)}]" if prompt else "")


def separate_speaker(audio_data):
    """Toy 2-stream 'separation' (position dependent), louder stream first like AudioProcessor.separate_speaker."""
    x = np.asarray(audio_data, dtype=np.float32)
    t = np.arange(x.size, dtype=np.float32) / max(x.size, 1)
    s1 = (x * (0.5 + t)).astype(np.float32)
    s2 = (np.tanh(3 * x) * 0.3 - 0.05 * np.sin(40 * t)).astype(np.float32)
    if meter_loudness(s1) < meter_loudness(s2):
        s1, s2 = s2, s1
    return s1, s2


class ToyEngine:
    """The engine protocol of targetdiarization_b200.streaming (meter_many / embed_many / separate_many) on the toys."""

    def __init__(self):
        self.calls = {"meter": 0, "embed": 0, "separate": 0}

    def meter_many(self, audios):
        self.calls["meter"] += 1
        return [meter_loudness(a) for a in audios]

    def embed_many(self, audios):
        self.calls["embed"] += 1
        return np.stack([embedding(a) for a in audios]) if len(audios) else np.zeros((0, 192), np.float32)

    def separate_many(self, audios):
        self.calls["separate"] += 1
        return [separate_speaker(a) for a in audios]


def scenario(seed=21, n_streams=4, n_steps=7):
    """chunks[step][stream], overlap[step][stream]: varied lengths (600 ms, 1.2 s, one under 0.4 s), one silent and one
    very weak chunk, a stream whose chunks alternate between two 'speakers'."""
    g = np.random.default_rng(seed)
    chunks, overlap = [], []
    for s in range(n_steps):
        row, orow = [], []
        for k in range(n_streams):
            n = [9600, 19200, 9600, 4800][(s + k) % 4] if not (s == 3 and k == 1) else 6000
            base = g.standard_normal(n).astype(np.float32)
            f = 1 + (k + (s % 2 if k == 2 else 0)) * 3
            x = np.convolve(base, np.ones(f, dtype=np.float32) / f, mode="same") * np.float32(0.05 * (1 + k))
            if s == 2 and k == 0:
                x[:] = 0.0
            if s == 4 and k == 3:
                x *= np.float32(0.01)
            if s == 5 and k == 1:
                x *= np.float32(0.02)
            row.append(x.astype(np.float32))
            orow.append(bool((s + 2 * k) % 3 == 0) and s > 0)
        chunks.append(row)
        overlap.append(orow)
    return chunks, overlap
