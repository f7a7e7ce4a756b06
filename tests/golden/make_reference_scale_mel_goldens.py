#!/usr/bin/env python
"""Golden vectors of MELInverter.scale_mel from the REAL reference (runs only where /root/reference is mounted).

``scale_mel`` (MBExWN_NVoc/mel_inverter.py:48-148) is plain NumPy / SciPy: the method is compiled unmodified from /root/reference
(the module itself imports TensorFlow, so it is extracted by AST) and called on a stand-in ``self`` carrying the attributes
``load_model`` sets (:212-239).  Cases: log-mel and linear-mel input, log / linear offsets and scales, an FFT size that differs
from the model's, a hop size that differs (re-interpolation), ``use_max_limit``.

Output: tests/golden/reference_scale_mel.npz (committed); tests/test_reference_source.py checks mbexwn_vocoder_b200.mel_inverter.
"""
import ast
import os
import sys
import types
from typing import Dict

import numpy as np
from scipy.interpolate import interp1d

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

MODEL = dict(hop_size=300, srate=24000, fft_size=2048, fmin=0, fmax=12000, mel_channels=80, lin_amp_scale=1, lin_amp_off=1.e-5,
             mel_amp_scale=1, use_max_limit=False)


def cases():
    rng = np.random.default_rng(11)
    mell = np.log(rng.uniform(1e-4, 1.0, size=(80, 37))).astype(np.float32)
    base = {"sr": 24000, "hoplen": 300, "winlen": 1200, "nfft": 2048, "fmin": 0, "fmax": 12000}
    return {
        "log": (dict(base, mell=mell.copy()), {}),
        "linear": (dict(base, mel=np.exp(mell)), {}),
        "log_offsets": (dict(base, mell=mell.copy() * 2.0 + 0.5, log_spec_offset=0.5, log_spec_scale=2.0,
                             lin_spec_offset=1e-5, lin_spec_scale=0.5), {}),
        "half_fft": (dict(base, mell=mell.copy(), nfft=1024), {}),
        "fft_size_key": ({k: v for k, v in dict(base, mell=mell.copy(), fft_size=512).items() if k != "nfft"}, {}),
        "slow_hop": (dict(base, mell=mell.copy(), hoplen=480), {}),
        "other_rate": (dict(base, mell=mell.copy(), sr=16000, hoplen=200, fmax=12000), {}),
        "model_scales": (dict(base, mell=mell.copy()), dict(lin_amp_scale=0.5, lin_amp_off=1e-3, mel_amp_scale=0.25, use_max_limit=True)),
    }


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    path = os.path.join(REF, "MBExWN_NVoc/mel_inverter.py")
    src = open(path).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "MELInverter"][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "scale_mel"][0]
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    code = "\n".join(ln[fn.col_offset:] for ln in lines)
    ns = {"np": np, "interp1d": interp1d, "sys": sys, "Dict": Dict, "log_to_db": 20 * np.log10(np.exp(1))}
    exec(compile(code, path + ":MELInverter.scale_mel", "exec"), ns)
    out = {}
    for tag, (cfg, model_extra) in cases().items():
        me = types.SimpleNamespace(**dict(MODEL, **model_extra))
        inp = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cfg.items()}
        res = ns["scale_mel"](me, inp)
        assert res.dtype == np.float32
        out[tag] = res
        print(tag, res.shape)
    dst = os.path.join(HERE, "reference_scale_mel.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
