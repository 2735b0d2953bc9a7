"""Host-side audio ingest of the drop-in wrapper (reference mellow/wrapper.py:141-179).

File decode -> resample to 32 kHz -> flatten channels -> tile-or-random-crop to exactly 320 000 samples.  The
reference uses ``torchaudio.load``; in this image torchaudio cannot decode (``torchcodec`` is absent), so PCM/float
RIFF files are read with the standard library / scipy and scaled like torchaudio does (int16 / 32768).
"""
import math
import random

import numpy as np
import torch


def read_wav(path, info_only=False):
    """-> (channels, samples) float32 in [-1, 1], sample_rate; info_only: -> (channels, samples per channel, rate)."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path, mmap=info_only)
    if data.ndim == 1:
        data = data[:, None]
    if info_only:
        return data.shape[1], data.shape[0], int(sr)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.T)), int(sr)


def read_audio(path, info_only=False):
    """What the reference gets from ``torchaudio.load`` (wrapper.py:144): (channels, samples) float32, sample rate.
    PCM / float RIFF files are decoded here; anything else (flac, mp3, ogg, compressed wav) goes to ``soundfile`` or
    ``torchaudio.load`` when this installation has a working decoder, and raises otherwise."""
    try:
        return read_wav(path, info_only)
    except Exception as wav_exc:
        audio = None
        try:
            import soundfile
            data, sr = soundfile.read(path, dtype="float32", always_2d=True)
            audio = torch.from_numpy(np.ascontiguousarray(data.T))
        except Exception:
            try:
                import torchaudio
                audio, sr = torchaudio.load(path)
            except Exception as exc:
                raise RuntimeError(f"cannot decode {path}: not a PCM/float WAV ({wav_exc}) and neither soundfile nor "
                                   f"torchaudio.load can read it here ({type(exc).__name__}: {exc})") from exc
        if info_only:
            return audio.shape[0], audio.shape[1], int(sr)
        return audio.to(torch.float32), int(sr)


def sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """The filter bank of torchaudio.transforms.Resample (sinc_interp_hann defaults), built in float64 and cast to
    float32 exactly like torchaudio.functional._get_sinc_resample_kernel does.
    -> (kernel [new, 2*width+orig] float32, width, orig, new) with orig/new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = int(math.ceil(lowpass_filter_width * orig / base))
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torchaudio evaluates the per-phase offset `arange(0, -new, -1) / new` on an int64 tensor, i.e. in float32
    # (torch's default dtype), before adding the float64 grid: reproduce that rounding
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)
    t = phase[:, None] + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2.0) ** 2
    t = t * math.pi
    scale = base / orig
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * window * scale
    return torch.from_numpy(kern.astype(np.float32)), width, orig, new


def resampled_length(n_in, orig, new):
    """Output length of torchaudio's resampler: ``torch.ceil(torch.as_tensor(new * length / orig))`` with orig / new
    reduced by their gcd (``_apply_sinc_resample_kernel``).  The float64 quotient is rounded to float32 by
    ``as_tensor`` BEFORE the ceil, which differs from ``math.ceil`` for ~1 % of the lengths (44.1 kHz -> 32 kHz,
    n = 300044: 217719 here and in torchaudio, 217720 in exact arithmetic)."""
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    return int(math.ceil(float(np.float32(new * int(n_in) / orig))))


def plan_fit(total, target, rng=random):
    """Tile-or-crop decision of reference wrapper.py:152-167 for a flattened clip of `total` samples: returns the
    start offset (0 when tiling; one `rng.randrange` draw when cropping, the same draw the reference makes)."""
    if target >= total:
        return 0
    return rng.randrange(total - target)


def load_audio_into_tensor(path, audio_duration, sample_rate_target, resample=True, rng=random):
    """Restates reference wrapper.py:141-168 (same branch conditions and the same ``random.randrange`` draw)."""
    audio, sr = read_audio(path)
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
