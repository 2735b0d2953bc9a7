"""Host-side audio ingest of the drop-in wrapper (reference mellow/wrapper.py:141-179).

File decode -> resample to 32 kHz -> flatten channels -> tile-or-random-crop to exactly 320 000 samples.  The
reference uses ``torchaudio.load``; in this image torchaudio cannot decode (``torchcodec`` is absent), so PCM/float
RIFF files are read with the standard library / scipy and scaled like torchaudio does (int16 / 32768).
"""
import math
import random

import numpy as np
import torch


def read_wav(path):
    """-> (channels, samples) float32 in [-1, 1], sample_rate."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)


def load_audio_into_tensor(path, audio_duration, sample_rate_target, resample=True, rng=random):
    """Restates reference wrapper.py:141-168 (same branch conditions and the same ``random.randrange`` draw)."""
    audio, sr = read_wav(path)
    if resample and sample_rate_target != sr:
        import torchaudio.transforms as T
        audio = T.Resample(sr, sample_rate_target)(audio)
    audio = audio.reshape(-1)                      # channels are concatenated, not mixed (wrapper.py:149)
    target = audio_duration * sample_rate_target
    if target >= audio.shape[0]:
        rep = int(math.ceil(target / audio.shape[0]))
        audio = audio.repeat(rep)[:target]
    else:
        start = rng.randrange(audio.shape[0] - target)
        audio = audio[start:start + target]
    return audio.to(torch.float32)
