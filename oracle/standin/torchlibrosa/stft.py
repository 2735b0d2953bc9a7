"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of torchlibrosa 0.1.0
``Spectrogram`` and ``LogmelFilterBank`` as used at reference
``mellow/model/htsat.py:647-653`` (constructed) and ``:864-865`` (called).

The parameter names (``stft.conv_real``, ``stft.conv_imag``, ``melW``) are the
ones the reference checkpoints carry (SURVEY.md section 8a').
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def hann_periodic(n: int) -> np.ndarray:
    """scipy/librosa ``get_window('hann', n, fftbins=True)``."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def dft_basis(n_fft: int):
    """Hann-windowed DFT rows: real[k,n]=w[n]cos(2 pi nk/N), imag[k,n]=-w[n]sin(2 pi nk/N)."""
    w = hann_periodic(n_fft)
    n = np.arange(n_fft, dtype=np.float64)[None, :]
    k = np.arange(n_fft // 2 + 1, dtype=np.float64)[:, None]
    ang = 2.0 * np.pi * ((n * k) % n_fft) / n_fft
    real = (np.cos(ang) * w[None, :]).astype(np.float32)
    imag = (-np.sin(ang) * w[None, :]).astype(np.float32)
    return real, imag


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> (n_mels, n_fft//2+1) float32."""
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


class STFT(nn.Module):
    def __init__(self, n_fft, hop_length, center, pad_mode):
        super().__init__()
        self.n_fft, self.hop_length, self.center, self.pad_mode = n_fft, hop_length, center, pad_mode
        out = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out, kernel_size=n_fft, stride=hop_length, bias=False)
        self.conv_imag = nn.Conv1d(1, out, kernel_size=n_fft, stride=hop_length, bias=False)
        real, imag = dft_basis(n_fft)
        self.conv_real.weight.data = torch.from_numpy(real)[:, None, :]
        self.conv_imag.weight.data = torch.from_numpy(imag)[:, None, :]
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        x = x[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)[:, None, :, :].transpose(2, 3)
        imag = self.conv_imag(x)[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                 pad_mode='reflect', power=2.0, freeze_parameters=True):
        super().__init__()
        assert window == 'hann' and win_length == n_fft
        self.power = power
        self.stft = STFT(n_fft, hop_length, center, pad_mode)

    def forward(self, x):
        real, imag = self.stft(x)
        spec = real ** 2 + imag ** 2
        if self.power != 2.0:
            spec = spec ** (self.power / 2.0)
        return spec


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True, ref=1.0,
                 amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log, self.ref, self.amin, self.top_db = is_log, ref, amin, top_db
        melW = slaney_mel(sr, n_fft, n_mels, fmin, fmax if fmax is not None else sr // 2).T
        self.melW = nn.Parameter(torch.from_numpy(np.ascontiguousarray(melW)), requires_grad=False)

    def forward(self, x):
        mel = torch.matmul(x, self.melW)
        if not self.is_log:
            return mel
        log_spec = 10.0 * torch.log10(torch.clamp(mel, min=self.amin, max=np.inf))
        log_spec = log_spec - 10.0 * np.log10(np.maximum(self.amin, self.ref))
        if self.top_db is not None:
            log_spec = torch.clamp(log_spec, min=log_spec.max().item() - self.top_db, max=np.inf)
        return log_spec
