"""CPU ORACLE of the analysis side (audio -> log-mel)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in NumPy, what ``MELInverter.generate_mel_from_snd`` (mel_inverter.py:156-182) runs:
``compute_mel_spectrogram_internal`` (vocoder/model/preprocess.py:417-560, band_limit=None) over ``calc_stft``
(sig_proc/spec/stft.py:14-96) and the librosa mel basis (preprocess.py:52-74), then ``scale_mel_spectrogram``
(preprocess.py:80-124) when ``do_post`` is set.

Pinning: ``calc_stft`` and the window generator are importable without TensorFlow, so ``stft_magnitude`` below is
checked against the REAL reference function (tests/golden/make_reference_analysis_goldens.py ->
tests/golden/reference_analysis.npz).  The mel basis needs librosa, which is absent: ``dsp_init.mel_filter_bank`` restates
its published algorithm and is anchored on its defining properties (tests/test_analysis.py) -- unpinned against librosa.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from mbexwn_vocoder_b200 import dsp_init


def stft_magnitude(x: np.ndarray, win_len: int, hop_len: int, fft_size: int, center: bool = True,
                   pad_mode: str = "reflect", dtype=np.float32) -> np.ndarray:
    """|STFT| of x (batch, time) -> (batch, frames, fft_size // 2 + 1); sig_proc/spec/stft.py:41-96 with do_mag=True.

    center: the signal is padded by (win_len // 2, win_len) and there are len // hop + 1 frames (:54-60)."""
    x = np.asarray(x)
    win = dsp_init.cosine_window("hann", win_len).astype(dtype)[np.newaxis]
    if center:
        num_frames = x.shape[-1] // hop_len + 1
        x = np.pad(x.astype(dtype, copy=False), ((0, 0), (win_len // 2, win_len)), mode=pad_mode)
    else:
        if x.shape[-1] < win_len:
            raise RuntimeError('calc_stft::error::cannot calculate STFT if signal is shorter than window')
        num_frames = (x.shape[-1] - win_len) // hop_len + 1
        x = x.astype(dtype, copy=False)
    res = np.empty((x.shape[0], num_frames, fft_size // 2 + 1), dtype=dtype)
    for ii in range(num_frames):
        start = ii * hop_len
        res[:, ii] = np.abs(np.fft.rfft(win * x[:, start:start + win_len], fft_size))
    return res


def scale_mel_spectrogram(mel: np.ndarray, preprocess_config: Dict) -> np.ndarray:
    """Forward branch of preprocess.py:80-108."""
    lin_amp_scale = preprocess_config["lin_amp_scale"] if preprocess_config.get("lin_amp_scale", 1) != 1 else 1
    lin_amp_off = preprocess_config["lin_amp_off"] if preprocess_config.get("lin_amp_off") is not None else 1.e-5
    mel_amp_scale = preprocess_config["mel_amp_scale"] if preprocess_config.get("mel_amp_scale", 1) != 1 else 1
    mel = np.array(mel) * lin_amp_scale
    if preprocess_config.get("use_max_limit"):
        return mel_amp_scale * np.log(np.fmax(mel, lin_amp_off)).astype(np.float32)
    return mel_amp_scale * np.log(mel + lin_amp_off).astype(np.float32)


def compute_mel_spectrogram(sound: np.ndarray, preprocess_config: Dict, do_post: bool = True,
                            dtype=np.float32, mel_basis: Optional[np.ndarray] = None) -> np.ndarray:
    """(batch, time) -> log-mel (batch, frames, mel_channels); preprocess.py:479-560 with band_limit=None, norm_mel off."""
    sound = np.asarray(sound)
    if sound.ndim == 1:
        sound = sound[np.newaxis, :]
    win_len = preprocess_config.get("win_size", preprocess_config["fft_size"])
    S = stft_magnitude(sound, win_len, preprocess_config["hop_size"], preprocess_config["fft_size"], dtype=dtype)
    if mel_basis is None:
        mel_basis = dsp_init.mel_filter_bank(preprocess_config["sample_rate"], preprocess_config["fft_size"],
                                             preprocess_config["mel_channels"], preprocess_config["fmin"],
                                             preprocess_config["fmax"], dtype=dtype)
    mel = np.dot(S, mel_basis.T)
    if do_post:
        if preprocess_config.get("norm_mel"):
            raise NotImplementedError("norm_mel is outside the restated path")
        return scale_mel_spectrogram(mel, preprocess_config)
    return np.log(np.fmax(mel, np.finfo(mel.dtype).eps))


def synthetic_audio(n_samples: int, utt_id: int = 0, sample_rate: int = 24000) -> np.ndarray:
    """Seeded test signal: a vibrato harmonic stack plus noise, peak about 0.5."""
    rng = np.random.default_rng(9000 + utt_id)
    t = np.arange(n_samples) / sample_rate
    f0 = 110.0 * (1 + utt_id % 5) * (1 + 0.03 * np.sin(2 * np.pi * 5.0 * t))
    ph = 2 * np.pi * np.cumsum(f0) / sample_rate
    x = sum(np.sin(k * ph) / k for k in range(1, 12))
    x = 0.25 * x / np.max(np.abs(x) + 1e-9) + 0.05 * rng.standard_normal(n_samples)
    return x.astype(np.float32)
