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


def fit_audio_unet(waveform: torch.FloatTensor, sampling_rate: int, fps: int, divisor_unet: int) -> FittedWaveform:
    """Zero-pad the waveform so that the coefficient sequence length is divisible by ``divisor_unet``
    (reference ``audio.py:42-75``)."""
    gcd = math.gcd(sampling_rate, fps)
    divisor_waveform = sampling_rate // gcd * divisor_unet

    waveform_len = waveform.shape[0]
    window_len = int(waveform_len / sampling_rate * fps)
    waveform_len_fit = math.ceil(waveform_len / divisor_waveform) * divisor_waveform

    if waveform_len_fit > waveform_len:
        tmp = torch.zeros(waveform_len_fit)
        tmp[:waveform_len] = waveform[:]
        waveform = tmp

    return FittedWaveform(waveform=waveform, window_size=window_len)
