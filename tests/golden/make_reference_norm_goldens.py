#!/usr/bin/env python
"""Golden vectors of the mel-derived RMS normaliser from the REAL reference source (runs only where /root/reference is mounted).

Same method as tests/golden/make_reference_excitation_goldens.py: class ``NormMelComponents`` (wavegen_1d.py:578-769) is compiled
unmodified from /root/reference -- its real constructor and ``normalize_inputs_by_rms`` run -- over NumPy float32 stand-ins for the
TensorFlow primitives it calls (exp, tensordot, matmul, overlap_and_add, conv1d with a stride, pow, log, maximum, concat, ones,
reshape, repeat, keras epsilon 1e-7).  ``get_stft_window`` is the reference's own function (sig_proc/spec/stft.py imports without
TensorFlow).  librosa is absent: ``librosa_mel_frequencies`` / ``get_mel_filter`` are taken from mbexwn_vocoder_b200.dsp_init (the
Slaney mel scale restated from its published definition) -- the band centres and the mel basis are therefore NOT pinned here, the
normaliser's algorithm (frame padding, Hann^2 overlap-add offsets, gain normalisation, smoothing iterations, re-scaling) is.

Output: tests/golden/reference_norm.npz (committed); tests/test_reference_source.py checks oracle/norm_mel.py against it.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_excitation_goldens as X                                      # noqa: E402
import make_reference_forward_goldens as FW                                        # noqa: E402

F32 = np.float32
CASES = {"default": {"normalize_rms_num_smooth_iters": 2},
         "one_iter_wide": {"normalize_rms_num_smooth_iters": 1, "normalize_smooth_win_scale": 2,
                           "normalize_smooth_with_squared_win": False, "max_norm_fact": 50.0, "normalize_compressor_exp": 0.8,
                           "use_max_limit": True},
         "pinv": {"normalize_rms_num_smooth_iters": 3, "normalize_use_pinv": True}}


def main():
    if not os.path.isdir(X.REF):
        print("reference not mounted; nothing to do")
        return 1
    sys.path.insert(0, X.REF)
    from MBExWN_NVoc.sig_proc.spec.stft import get_stft_window                      # the reference's own window function
    from mbexwn_vocoder_b200 import dsp_init, get_config_file
    from mbexwn_vocoder_b200.config import read_config
    from oracle.forward import synthetic_mel

    tf = X.make_tf({})
    tf.Module = type("Module", (), {"__init__": lambda self, *a, **k: None})
    tf.exp = lambda x: np.exp(x).astype(F32)
    tf.square = lambda x: np.square(x).astype(F32)
    tf.tensordot = lambda a, b, axes: np.tensordot(a, b, axes=axes).astype(F32)
    tf.ones = lambda shape, dtype=F32: np.ones(tuple(int(s) for s in shape), dtype=dtype)
    tf.repeat = lambda x, n, axis: np.repeat(x, int(n), axis=axis)
    tf.reduce_mean = lambda x, axis=None, keepdims=False: np.mean(x, axis=axis, keepdims=keepdims, dtype=F32)
    tf.linalg = types.SimpleNamespace(matmul=lambda a, b: np.matmul(a, b).astype(F32))
    tf.signal = types.SimpleNamespace(overlap_and_add=FW.overlap_and_add)
    tf.keras.backend = types.SimpleNamespace(epsilon=lambda: 1e-7)
    tf.math.pow = lambda x, p: np.power(x, F32(p)).astype(F32)
    tf.nn.conv1d = lambda x, f, stride=1, padding="VALID", data_format="NWC": X.conv1d(
        x, f, None, padding, 1, stride[1] if isinstance(stride, (list, tuple)) else stride)
    ns = {"tf": tf, "np": np, "copy": __import__("copy"), "get_stft_window": get_stft_window,
          "librosa_mel_frequencies": lambda n_mels, fmin, fmax, htk=False: dsp_init.mel_frequencies(n_mels, fmin, fmax),
          "get_mel_filter": lambda sr, n_fft, n_mels, fmin, fmax, dtype="float32": dsp_init.mel_filter_bank(sr, n_fft, n_mels, fmin, fmax),
          "LinInterpLayer": None}
    path = os.path.join(X.REF, "MBExWN_NVoc/vocoder/model/wavegen_1d.py")
    exec(compile(X._segments(path, {"NormMelComponents"})["NormMelComponents"], path + ":NormMelComponents", "exec"), ns)

    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    out = {}
    T = 14
    mell = np.stack([synthetic_mel(T, 70 + i) for i in range(2)]).astype(F32)
    out["mell"] = mell
    for tag, kw in CASES.items():
        nm = ns["NormMelComponents"](pc, **kw)
        for length, suffix in ((T * pc["hop_size"], ""), (T * pc["hop_size"] + 170, "_long")):
            _, out_mell, up = nm.normalize_inputs_by_rms(None, mell, synth_length=length)
            assert out_mell.shape == mell.shape and up.shape == (2, length, 1), (out_mell.shape, up.shape)
            out[f"{tag}{suffix}_mell"], out[f"{tag}{suffix}_rms"] = out_mell.astype(F32), up[:, :, 0].astype(F32)
        print(tag, "rms range", float(up.min()), float(up.max()))
    dst = os.path.join(HERE, "reference_norm.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
