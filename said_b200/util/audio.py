"""Audio helpers with the reference's interface (``said/util/audio.py:12-75``)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class FittedWaveform:
    """Fitted waveform using the window (reference ``audio.py:12-17``)"""

    waveform: torch.FloatTensor
    window_size: int


def _read_wav(audio_path: str):
    """(channels, samples) float32 in [-1, 1] and the sampling rate.  torchaudio.load is used when its
    backend is usable; otherwise PCM/float WAV files are read with scipy (torchaudio >= 2.9 needs
    torchcodec, which offline images do not ship)."""
    try:
        import torchaudio

        wav, sr = torchaudio.load(audio_path)
        return wav.to(torch.float32), int(sr)
    except Exception:
        from scipy.io import wavfile

        sr, data = wavfile.read(audio_path)
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":  # 8-bit PCM
            data = (data.astype(np.float32) - 128.0) / 128.0
        else:
            data = data.astype(np.float32)
        if data.ndim == 1:
            data = data[None]
        else:
            data = data.T
        return torch.from_numpy(np.ascontiguousarray(data)), int(sr)


def load_audio(audio_path: str, sampling_rate: int) -> torch.FloatTensor:
    """Load the audio file as a mono waveform (T_a,) at ``sampling_rate`` (reference ``audio.py:20-39``)."""
    waveform, sr = _read_wav(audio_path)
    if sr != sampling_rate:
        import torchaudio

        waveform = torchaudio.functional.resample(waveform, sr, sampling_rate)
    return torch.mean(waveform, dim=0)


def load_audio_device(audio_path: str, sampling_rate: int, device) -> torch.FloatTensor:
    """``load_audio`` with the resampling and the mono mix on the GPU: the file is decoded on the host (WAV
    container parsing is not device work), uploaded as stored, and resampled / averaged over channels by ``said_resample_mono``
    (torchaudio's algorithm, its default parameters).  Returns the mono waveform (T_a,) on ``device``."""
    from .._lib import Engine

    waveform, sr = _read_wav(audio_path)
    eng = _engine_for(device)
    wave_dev = waveform.to(eng.device, torch.float32).contiguous()
    if sr != sampling_rate:
        return eng.resample_mono(wave_dev, sr, sampling_rate)
    return wave_dev.mean(dim=0)


_ENGINES = {}


def _engine_for(device):
    from .._lib import Engine

    d = torch.device(device)
    key = d.index if d.index is not None else torch.cuda.current_device()
    if key not in _ENGINES:
        _ENGINES[key] = Engine(torch.device("cuda", key))
    return _ENGINES[key]


def fit_audio_unet(waveform: torch.FloatTensor, sampling_rate: int, fps: int, divisor_unet: int) -> FittedWaveform:
    """Zero-pad the mono waveform (host or device tensor) up to the next multiple of the hop that makes the coefficient sequence
    length divisible by ``divisor_unet``; ``window_size`` is the frame count of the UNPADDED audio (reference ``audio.py:42-75``)."""
    hop = sampling_rate // math.gcd(sampling_rate, fps) * divisor_unet      # samples per `divisor_unet` coefficient frames
    n = int(waveform.shape[0])
    deficit = -n % hop
    fitted = torch.nn.functional.pad(waveform, (0, deficit)) if deficit else waveform
    return FittedWaveform(waveform=fitted, window_size=int(n / sampling_rate * fps))
