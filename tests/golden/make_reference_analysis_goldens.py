#!/usr/bin/env python
"""Golden vectors of the analysis side from the REAL reference code (runs only where /root/reference is mounted).

``MBExWN_NVoc.sig_proc.spec.stft.calc_stft`` and ``MBExWN_NVoc.sig_proc.Mwindows.window`` are plain NumPy and import without
TensorFlow; they are executed as they are on a seeded signal.  Output: tests/golden/reference_analysis.npz (committed):
the signal, the reference's Hann window and |STFT| exactly as ``compute_mel_spectrogram_internal`` asks for it
(preprocess.py:486-489: win 1200, hop 300, fft 2048, center, reflect, do_mag, float32).
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    sys.path.insert(0, REF)
    from MBExWN_NVoc.sig_proc.spec.stft import calc_stft
    from MBExWN_NVoc.sig_proc.Mwindows import window
    from oracle.analysis import synthetic_audio
    x = np.stack([synthetic_audio(4000, 0), synthetic_audio(4000, 3)])
    S = calc_stft(x, win_len=1200, hop_len=300, fft_size=2048, win_type='hann', center=True, pad_mode="reflect",
                  do_mag=True, axis=-1, dtype=np.dtype('float32'))
    short = synthetic_audio(700, 1)[None]                  # shorter than the window: the reflect pad wraps more than once
    S_short = calc_stft(short, win_len=1200, hop_len=300, fft_size=2048, win_type='hann', center=True,
                        pad_mode="reflect", do_mag=True, axis=-1, dtype=np.dtype('float32'))
    out = os.path.join(HERE, "reference_analysis.npz")
    np.savez_compressed(out, audio=x, stft_mag=S, audio_short=short, stft_mag_short=S_short,
                        hann1200=window("hann", 1200), hann7=window("hann", 7), hamming8=window("hamming", 8))
    print("wrote", out, S.shape, S.dtype, S_short.shape)
    return 0


if __name__ == "__main__":
    sys.exit(main())
