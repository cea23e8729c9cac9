"""Host-side constant tables of the Kaldi fbank front-end (window, FFT twiddles, mel filterbank).

The published Kaldi / torchaudio definitions (torchaudio/compliance/kaldi.py:154-217 `_get_window` povey,
:436-511 `get_mel_banks`) restated with numpy in float64 and rounded once to float32; the kernel
(csrc/kernels_fbank.cuh) does the rest on the GPU."""
import math

import numpy as np

SAMPLE_RATE = 16000
WIN = 400        # 25 ms
SHIFT = 160      # 10 ms
NFFT = 512
NMEL = 80
NBIN = NFFT // 2 + 1
LOW_FREQ = 20.0


def num_frames(n_samples):
    """snip_edges=True: 1 + (T - 400) // 160 (0 if shorter than one window)."""
    return 0 if n_samples < WIN else 1 + (n_samples - WIN) // SHIFT


def povey_window():
    n = np.arange(WIN, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / (WIN - 1))
    return (hann ** 0.85).astype(np.float32)


def twiddles():
    k = np.arange(NFFT // 2, dtype=np.float64)
    ang = -2.0 * math.pi * k / NFFT
    return np.stack((np.cos(ang), np.sin(ang)), axis=-1).astype(np.float32)


def _mel(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def mel_banks():
    """[80, 257] triangular filters on the mel scale (Kaldi get_mel_banks; last column is the zero pad),
    plus first/last non-zero bin per filter.  Evaluated with float32 torch ops in the order torchaudio uses
    (kaldi.py:436-511), so the table is bit-identical to the one the reference pipeline multiplies by."""
    import torch
    bin_width = SAMPLE_RATE / NFFT
    mel_lo = 1127.0 * math.log(1.0 + LOW_FREQ / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + 0.5 * SAMPLE_RATE / 700.0)
    delta = (mel_hi - mel_lo) / (NMEL + 1)
    b = torch.arange(NMEL).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (bin_width * torch.arange(NFFT // 2)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = torch.max(torch.zeros(1), torch.min(up, down)).numpy()
    full = np.zeros((NMEL, NBIN), dtype=np.float32)
    full[:, :NFFT // 2] = w
    lo = np.zeros(NMEL, dtype=np.int32)
    hi = np.zeros(NMEL, dtype=np.int32)
    for i in range(NMEL):
        nz = np.nonzero(full[i])[0]
        lo[i], hi[i] = (nz[0], nz[-1]) if len(nz) else (0, -1)
    return full, lo, hi
