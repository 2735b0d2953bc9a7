"""Closed forms of the frozen front-end tensors a Mellow checkpoint carries.

The reference stores the STFT basis and the mel matrix as frozen parameters
(torchlibrosa ``Spectrogram`` / ``LogmelFilterBank`` built at reference
``mellow/model/htsat.py:647-653``; keys in SURVEY.md section 8a').  The FFT
front-end kernel is only valid when the checkpoint's basis *is* the
Hann-windowed DFT, so the weight packer regenerates the closed form here and
compares (``weights.check_frontend_basis``).  The same closed forms seed the
synthetic checkpoint.
"""
import numpy as np

from . import schema as S


def periodic_hann(n=S.N_FFT):
    i = np.arange(n, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * i / n))


def windowed_dft(n=S.N_FFT):
    """(real, imag), each (n/2+1, n) float32: w[j]*cos(2*pi*j*k/n), -w[j]*sin(2*pi*j*k/n)."""
    w = periodic_hann(n)
    j = np.arange(n, dtype=np.int64)[None, :]
    k = np.arange(n // 2 + 1, dtype=np.int64)[:, None]
    phase = 2.0 * np.pi * ((j * k) % n).astype(np.float64) / n
    return (w * np.cos(phase)).astype(np.float32), (-w * np.sin(phase)).astype(np.float32)


def _mel_of_hz(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    knee_hz, knee_mel, step = 1000.0, 15.0, np.log(6.4) / 27.0
    return np.where(f >= knee_hz, knee_mel + np.log(np.maximum(f, 1e-30) / knee_hz) / step, lin)


def _hz_of_mel(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * 200.0 / 3.0
    knee_hz, knee_mel, step = 1000.0, 15.0, np.log(6.4) / 27.0
    return np.where(m >= knee_mel, knee_hz * np.exp(step * (m - knee_mel)), lin)


def slaney_mel_matrix(sr=S.SAMPLE_RATE, n_fft=S.N_FFT, n_mels=S.N_MELS, fmin=S.FMIN, fmax=S.FMAX):
    """(n_fft/2+1, n_mels) float32 triangular Slaney-normalised filterbank (librosa.filters.mel(...).T)."""
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = _hz_of_mel(np.linspace(_mel_of_hz(fmin), _mel_of_hz(fmax), n_mels + 2))
    out = np.zeros((n_fft // 2 + 1, n_mels), dtype=np.float64)
    for m in range(n_mels):
        lo, mid, hi = edges[m], edges[m + 1], edges[m + 2]
        rising = (bins - lo) / (mid - lo)
        falling = (hi - bins) / (hi - mid)
        out[:, m] = np.maximum(0.0, np.minimum(rising, falling)) * (2.0 / (hi - lo))
    return out.astype(np.float32)


def relative_position_index(window=S.WINDOW):
    """(w*w, w*w) int64 index into the (2w-1)^2 bias table (reference htsat.py:281-290)."""
    r = np.arange(window)
    ch, cw = np.meshgrid(r, r, indexing="ij")
    ch, cw = ch.reshape(-1), cw.reshape(-1)
    dh = ch[:, None] - ch[None, :] + window - 1
    dw = cw[:, None] - cw[None, :] + window - 1
    return (dh * (2 * window - 1) + dw).astype(np.int64)


def shift_region_id(res, window=S.WINDOW, shift=S.WINDOW // 2):
    """(res,res) region labels 0..8 used to build the shifted-window mask (reference htsat.py:391-403)."""
    lab = np.zeros((res, res), dtype=np.int64)
    def band(x):
        return np.where(x < res - window, 0, np.where(x < res - shift, 1, 2))
    b = band(np.arange(res))
    return b[:, None] * 3 + b[None, :]


def shifted_window_mask(res, window=S.WINDOW, shift=S.WINDOW // 2):
    """(nW, w*w, w*w) float32 with 0 / -100 (reference htsat.py:405-408)."""
    lab = shift_region_id(res, window, shift)
    nw = res // window
    lab = lab.reshape(nw, window, nw, window).transpose(0, 2, 1, 3).reshape(nw * nw, window * window)
    diff = lab[:, None, :] - lab[:, :, None]
    return np.where(diff != 0, -100.0, 0.0).astype(np.float32)
