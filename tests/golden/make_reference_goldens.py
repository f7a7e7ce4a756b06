#!/usr/bin/env python
"""Generate golden vectors from the REAL reference code (runs only where /root/reference is mounted).

The reference's hot path needs TensorFlow, which is not installed, but its init-time DSP is plain
NumPy/SciPy.  This script executes those reference functions *as they are* (source is read from
/root/reference at run time, never copied into this repo):

* ``MBExWN_NVoc.glottis.FglotspecLF`` / ``FglotLFsynthparams`` are imported directly;
* ``get_pulse_lowpass_kaiser``, ``get_LFpulse``, ``PulseWaveTable.create_normed_pulse``
  (tf_wavetable.py) and ``_design_prototype_filter`` (tf_preprocess.py) are extracted from the
  module source by AST (the modules import tensorflow at top level so they cannot be imported) and
  executed unmodified, with three NumPy-2/SciPy-1.18 shims: ``np.int``, ``np.float`` and
  ``scipy.signal.kaiser``;
* the table loop of ``PulseWaveTable.__init__`` (tf_wavetable.py:244-291) is driven from here with the
  extracted functions (that constructor itself calls tf.constant/tf.concat).

Output: tests/golden/reference_init_dsp.npz (committed).  tests/test_dsp_init.py checks
mbexwn_vocoder_b200.dsp_init against it on any machine.
"""
import ast
import os
import sys
import types

import numpy as np
import scipy.signal as ss
import scipy.signal.windows

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _extract(path, names, class_name=None):
    """Return source segments of top-level (or class-level) function defs called `names`."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    bodies = tree.body
    if class_name is not None:
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == class_name:
                bodies = node.body
    for node in bodies:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            seg = ast.get_source_segment(src, node)
            # drop decorators such as @staticmethod, dedent class-level defs
            lines = seg.splitlines()
            indent = len(lines[0]) - len(lines[0].lstrip())
            col = node.col_offset
            lines = [lines[0]] + [ln[col:] if len(ln) >= col else ln for ln in lines[1:]]
            out[node.name] = "\n".join(lines)
            del indent
    return out


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    sys.path.insert(0, REF)
    # the package __init__ is importable (no TF); glottis is plain NumPy/SciPy
    from MBExWN_NVoc.glottis.FglotspecLF import FglotspecLF
    from MBExWN_NVoc.glottis.FglotLFsynthparams import FglotLFsynthparams

    np_shim = types.ModuleType("np_shim")
    np_shim.__dict__.update(np.__dict__)
    np_shim.int = int
    np_shim.float = float
    ss_shim = types.ModuleType("ss_shim")
    ss_shim.__dict__.update(ss.__dict__)
    ss_shim.kaiser = scipy.signal.windows.kaiser
    ns = {"np": np_shim, "ss": ss_shim, "FglotspecLF": FglotspecLF, "print": lambda *a, **k: None}

    wt_path = os.path.join(REF, "MBExWN_NVoc/vocoder/model/tf_wavetable.py")
    pp_path = os.path.join(REF, "MBExWN_NVoc/vocoder/model/tf_preprocess.py")
    for name, code in _extract(wt_path, {"get_pulse_lowpass_kaiser", "get_min_phase_spectrum", "get_LFpulse"}).items():
        exec(compile(code, wt_path + ":" + name, "exec"), ns)
    for name, code in _extract(wt_path, {"create_normed_pulse"}, class_name="PulseWaveTable").items():
        exec(compile(code, wt_path + ":" + name, "exec"), ns)
    tf_stub = types.SimpleNamespace(Variable=type("Variable", (), {}))
    ns["tf"] = tf_stub
    for name, code in _extract(pp_path, {"_design_prototype_filter"}).items():
        exec(compile(code, pp_path + ":" + name, "exec"), ns)

    create_normed_pulse = ns["create_normed_pulse"]
    out = {}

    # ---- LF synthesis parameters and spectra -------------------------------------------------
    lf_cases = [(0.5, 0.8, 0.025), (0.6, 0.7, 0.0), (0.4, 0.9, 0.1), (0.7, 0.75, 0.299), (0.9995, 0.8, 0.0002)]
    f = np.arange(0, 65) * 0.37
    for i, (oq, am, ta) in enumerate(lf_cases):
        a, e, t = FglotLFsynthparams(oq, am, ta)
        out[f"lfpar_{i}"] = np.array([oq, am, ta, a, e, t])
        out[f"lfspec_flow_{i}"] = FglotspecLF(f, oq=oq, am=am, ta=ta, get_derivative=False, orig=0)[0]
        out[f"lfspec_deriv_{i}"] = FglotspecLF(f, oq=oq, am=am, ta=ta, get_derivative=True, orig=0)[0]
    out["lf_freqs"] = f

    # ---- wavetable banks (driver = tf_wavetable.py:244-291) ----------------------------------
    wt_cases = {"sp": dict(sample_rate=8000.0, nominalF0=60.0, maxF0=550.0),
                "vo": dict(sample_rate=8000.0, nominalF0=60.0, maxF0=1400.0),
                "alt": dict(sample_rate=8000.0, nominalF0=80.0, maxF0=800.0)}
    for tag, c in wt_cases.items():
        Oq, am, rta, grid_factor, over = 0.5, 0.8, 0.05, 1.25, 2
        _, nominal = create_normed_pulse(Oq, target_nominalF0=c["nominalF0"], nominalBandWidth=0.5 / grid_factor,
                                         sample_rate=c["sample_rate"], am=am, rta=rta, use_radiation=False,
                                         bandWidthReductionFactor=c["maxF0"] / c["nominalF0"],
                                         wt_oversampling=over, return_nominal_f0=True, quiet=True,
                                         use_sinusoid=False, use_white_pulse=False)
        n_grid = int(np.ceil(np.log(c["maxF0"] / nominal) / np.log(grid_factor)))
        cols, grid = [], []
        for ir in range(n_grid + 1):
            rs = grid_factor ** ir if ir > 0 else 1
            wt = create_normed_pulse(Oq, target_nominalF0=nominal, nominalBandWidth=0.5,
                                     sample_rate=c["sample_rate"], am=am, rta=rta, use_radiation=False,
                                     bandWidthReductionFactor=rs, wt_oversampling=over, use_sinusoid=False,
                                     quiet=True, use_white_pulse=False).astype(np.float32)
            grid.append(nominal * rs)
            cols.append(np.concatenate([wt, wt[0:1]], axis=0)[:, np.newaxis])
        norm = -np.min([cols])
        out[f"wt_{tag}_tables"] = np.concatenate([w / norm for w in cols], axis=1).astype(np.float32)
        out[f"wt_{tag}_grid"] = np.asarray(grid)
        out[f"wt_{tag}_cfg"] = np.array([c["sample_rate"], c["nominalF0"], c["maxF0"], nominal])

    # ---- PQMF prototype ----------------------------------------------------------------------
    out["pqmf_proto_240"] = ns["_design_prototype_filter"](240, 0.0377, 9.0)
    out["pqmf_proto_62"] = ns["_design_prototype_filter"](62, 0.15, 9.0)

    # ---- stft window helper (sig_proc is NumPy-only) -------------------------------------------
    np.savez_compressed(os.path.join(HERE, "reference_init_dsp.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_init_dsp.npz"), {k: np.shape(v) for k, v in out.items() if k.startswith("wt_")})
    return 0


if __name__ == "__main__":
    sys.exit(main())
