#!/usr/bin/env python
"""Golden vectors of the pulse generator from the REAL reference source (runs only where /root/reference is mounted).

``PulseWaveTable.call``, ``stable_cumsum_and_wrap`` and ``_linear_lookup`` (tf_wavetable.py:429-638) are TensorFlow code, and
TensorFlow is not installable here.  This script still executes those three methods *unmodified* -- their source is read from
/root/reference at run time and compiled as it stands, never copied into the repo -- against a small stand-in for the ``tf``
module in which every TensorFlow primitive they call (reshape, cumsum, pad, %, floor, cast, gather, concat, reduce_sum,
maximum, minimum, log, abs, sin, range, constant) is the NumPy operation of the same meaning in float32.  What this pins is the
reference's *algorithm as written*: padding to whole chunks, the chunk reshape, where the wrap is applied, the shifted offsets,
index = floor(phase * n_period), the two-tap lookup and the cross-fade between the band-limited tables.  What it cannot pin
is TensorFlow's own arithmetic inside a primitive; the stand-ins state the assumed meaning:

* ``tf.cumsum``: sequential float32 accumulation along the axis (Eigen scan on CPU) = ``np.cumsum(dtype=float32)``;
* ``%`` on tensors: floor-mod = NumPy ``%``;  ``tf.cast(float -> int32)``: truncation (the argument is already floored);
* ``tf.gather(params, indices, axis=0, batch_dims=1)``: array_ops.gather takes the axis == 0 branch (``sparse_read`` /
  ``gather_v2`` *without* batch_dims), i.e. plain ``params[indices]``;
* ``tf.math.log`` / ``tf.sin``: NumPy float32 (may differ from Eigen in the last ulp -- the float outputs are compared with a
  tolerance, the integer index and the wrapped phase bit for bit).

The wavetable bank fed to the methods is the one ``reference_init_dsp.npz`` already pins to the reference's own construction code.

Output: tests/golden/reference_pulse.npz (committed); tests/test_reference_source.py checks the oracle against it on any machine,
tests/test_gpu_parity.py the CUDA kernels (index and phase bit-exact).
"""
import ast
import os
import sys
import types
from typing import List, Union

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _extract(path, names, class_name=None):
    src = open(path).read()
    tree = ast.parse(src)
    bodies = tree.body
    if class_name is not None:
        bodies = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0].body
    out = {}
    for node in bodies:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            lines = ast.get_source_segment(src, node).splitlines()          # starts at `def`: decorators are not part of it
            col = node.col_offset
            out[node.name] = "\n".join([lines[0]] + [ln[col:] if len(ln) >= col else ln for ln in lines[1:]])
    return out


def make_tf_stand_in(gather_log):
    f32 = np.float32

    def _f(x):
        return np.asarray(x, dtype=f32) if not isinstance(x, np.ndarray) else x

    def reshape(x, shape):
        return np.reshape(x, [int(s) for s in shape])

    def gather(params, indices, axis=0, batch_dims=0):
        assert axis == 0                                  # array_ops.gather: axis 0 ignores batch_dims
        gather_log.append(np.array(indices))
        return np.asarray(params)[indices]

    tf = types.SimpleNamespace(
        newaxis=None, float32=np.float32, int32=np.int32, int64=np.int64, bool=np.bool_, Tensor=np.ndarray,
        reshape=reshape, gather=gather,
        cumsum=lambda x, axis=0: np.cumsum(x, axis=axis, dtype=x.dtype),
        pad=lambda x, paddings, **kw: np.pad(x, paddings),
        constant=lambda v, dtype=np.float32: np.asarray(v, dtype=dtype)[()],
        cast=lambda x, dtype: np.asarray(x).astype(dtype) if isinstance(x, np.ndarray) else dtype(x),
        range=lambda n: np.arange(n, dtype=np.int32),
        floor=lambda x: np.floor(_f(x)), abs=lambda x: np.abs(_f(x)), sin=lambda x: np.sin(_f(x)),
        maximum=lambda a, b: np.maximum(_f(a), _f(b)), minimum=lambda a, b: np.minimum(_f(a), _f(b)),
        reduce_sum=lambda x, axis=None: np.sum(x, axis=axis, dtype=x.dtype),
        concat=lambda xs, axis=0: np.concatenate(list(xs), axis=axis),
        math=types.SimpleNamespace(log=lambda x: np.log(_f(x))))
    return tf


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    wt_path = os.path.join(REF, "MBExWN_NVoc/vocoder/model/tf_wavetable.py")
    gather_log = []
    tf = make_tf_stand_in(gather_log)
    ns = {"tf": tf, "np": np, "Union": Union, "List": List}
    for name, code in _extract(wt_path, {"pad_axis"}).items():
        exec(compile(code, wt_path + ":" + name, "exec"), ns)
    methods = _extract(wt_path, {"stable_cumsum_and_wrap", "call", "_linear_lookup"}, class_name="PulseWaveTable")
    assert sorted(methods) == ["_linear_lookup", "call", "stable_cumsum_and_wrap"]
    for name, code in methods.items():
        exec(compile(code, wt_path + ":" + name, "exec"), ns)

    gold = np.load(os.path.join(HERE, "reference_init_dsp.npz"))
    out = {}
    rng = np.random.default_rng(7)
    for tag, grid_factor in (("sp", 1.25), ("vo", 1.25)):
        tables, grid, cfg = gold[f"wt_{tag}_tables"], gold[f"wt_{tag}_grid"], gold[f"wt_{tag}_cfg"]
        sample_rate, nominal = float(cfg[0]), cfg[3]
        for subharm in (0, 2):
            obj = types.SimpleNamespace(
                sample_rate=sample_rate, use_sinusoid_as_fun=False, add_subharm_chans=subharm, pulse_sync_gain_avg=False,
                no_interp=False, wavetables=tables, n_period=int(tables.shape[0] - 1), nominalF0=nominal,
                # tf_wavetable.py:283-284, :305
                minTranspositionFactorInGrid=tf.constant(np.min(grid) / nominal, tf.float32),
                maxTranspositionFactorInGrid=tf.constant(np.max(grid) / nominal, tf.float32),
                grid_f0_diff_norm_factor=1. / tf.math.log(grid_factor))
            obj.stable_cumsum_and_wrap = types.MethodType(ns["stable_cumsum_and_wrap"], obj)
            obj._linear_lookup = types.MethodType(ns["_linear_lookup"], obj)
            # three F0 contours per length: constant, exponential sweep over more than the model's range, random walk;
            # lengths on and off the 1000-sample chunk grid (multiples of 100 = whole mel frames of 100 pulse samples)
            lo, hi = (45.0, 700.0) if tag == "sp" else (45.0, 1400.0)
            # (the sub-harmonic channels are kept for one short case only: fixture size)
            for n in ((2300, 1000, 700) if subharm == 0 else ((700,) if tag == "sp" else ())):
                t = np.arange(n, dtype=np.float64) / n
                f0 = np.stack([np.full(n, 123.456), lo * (hi / lo) ** t,
                               np.clip(200.0 * np.exp(np.cumsum(rng.normal(0, 0.01, n))), lo, hi)]).astype(np.float32)
                gather_log.clear()
                phase = obj.stable_cumsum_and_wrap(f0 / obj.sample_rate)
                audio = ns["call"](obj, f0)
                assert len(gather_log) == 1 and gather_log[0].shape == (3, n, 2)
                key = f"{tag}_s{subharm}_{n}"
                out[key + "_f0"], out[key + "_phase"], out[key + "_audio"] = f0, phase, audio
                out[key + "_index"] = gather_log[0][:, :, 0].astype(np.int32)
                assert phase.dtype == np.float32 and audio.dtype == np.float32 and audio.shape == (3, n, 1 + subharm)
    path = os.path.join(HERE, "reference_pulse.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays,", os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
