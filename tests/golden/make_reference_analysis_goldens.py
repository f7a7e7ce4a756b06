#!/usr/bin/env python
"""Golden vectors of the analysis side from the REAL reference code (runs only where /root/reference is mounted).

``MBExWN_NVoc.sig_proc.spec.stft.calc_stft`` and ``MBExWN_NVoc.sig_proc.Mwindows.window`` are plain NumPy and import without
TensorFlow; they are executed as they are on a seeded signal.  Output: tests/golden/reference_analysis.npz (committed):
the signal, the reference's Hann window and |STFT| exactly as ``compute_mel_spectrogram_internal`` asks for it
(preprocess.py:486-489: win 1200, hop 300, fft 2048, center, reflect, do_mag, float32).

``compute_mel_spectrogram_internal`` and ``scale_mel_spectrogram`` (preprocess.py:81-113, :417-572) are NumPy as well, but their module
imports -- and calls -- librosa at import time: the two functions are AST-extracted and run unmodified with ``get_mel_filter`` bound to
this package's Slaney mel basis (dsp_init.mel_filter_bank; librosa is absent, so the basis itself is not pinned, the framing, the
projection and the log / scale post-processing are).  Stored as ``mell_post`` / ``mell_nopost`` / ``mell_post_scaled``.
"""
import ast
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
    # the caller of calc_stft: compute_mel_spectrogram_internal + scale_mel_spectrogram from the reference's source
    from MBExWN_NVoc.sig_proc.spec.stft import get_stft_window
    from mbexwn_vocoder_b200 import dsp_init, get_config_file
    from mbexwn_vocoder_b200.config import read_config
    pp_path = os.path.join(REF, "MBExWN_NVoc/vocoder/model/preprocess.py")
    src = open(pp_path).read()
    ns = {"np": np, "sys": sys, "calc_stft": calc_stft, "get_stft_window": get_stft_window, "have_STFT": False,
          "get_mel_filter": lambda sr, n_fft, n_mels, fmin, fmax, dtype=np.float32: dsp_init.mel_filter_bank(sr, n_fft, n_mels, fmin, fmax, dtype=dtype)}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("scale_mel_spectrogram", "compute_mel_spectrogram_internal"):
            exec(compile("\n".join(src.splitlines()[node.lineno - 1:node.end_lineno]), pp_path + ":" + node.name, "exec"), ns)
    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    mell_post, rate = ns["compute_mel_spectrogram_internal"](x, pc)
    mell_nopost, _ = ns["compute_mel_spectrogram_internal"](x, pc, do_post=False)
    pc2 = dict(pc, lin_amp_scale=0.5, lin_amp_off=1e-3, mel_amp_scale=0.25, use_max_limit=True)
    mell_scaled, _ = ns["compute_mel_spectrogram_internal"](x, pc2)
    assert rate == pc["sample_rate"] / pc["hop_size"]
    out = os.path.join(HERE, "reference_analysis.npz")
    np.savez_compressed(out, audio=x, stft_mag=S, audio_short=short, stft_mag_short=S_short,
                        mell_post=mell_post, mell_nopost=mell_nopost, mell_post_scaled=mell_scaled,
                        hann1200=window("hann", 1200), hann7=window("hann", 7), hamming8=window("hamming", 8))
    print("wrote", out, S.shape, S.dtype, S_short.shape)
    return 0


if __name__ == "__main__":
    sys.exit(main())
