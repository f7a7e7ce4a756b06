#!/usr/bin/env python
"""Golden vectors of the excitation branch (pulse train -> WaveNet block -> post net -> PQMF synthesis) from the REAL reference
source (runs only where /root/reference is mounted).

TensorFlow is not installable here, so the reference cannot run as a program.  Its *source* still can: this script compiles, from
/root/reference and unmodified (nothing is copied into the repo),

* ``MBExWN.generate_excitation``                                        custom_pulsed_generator.py:886-925
* class ``WaveNetAE`` (``call``) and ``WaveNetAEBlock`` (``call``)       custom_AE_layers.py:114-346, :457-575
* classes ``TF2C_Conv1DWeightNorm`` / ``TF2C_Conv1DUpDownSample``        conv_layers.py:21-261  (``call``: weight-norm fold, sub-pixel unfold)
* class ``TF2C_LinInterpLayer`` + its weight initialiser               support_layers.py:19-121
* class ``TFPQMF`` (real constructor, ``synthesis``) + prototype design tf_preprocess.py:30-226
* class ``PulseWaveTable`` (``call`` / ``stable_cumsum_and_wrap`` / ``_linear_lookup``)   tf_wavetable.py:429-638

and executes them over a stand-in for the ``tf`` module in which each TensorFlow / Keras *primitive* is a NumPy float32 function
written from its documented meaning (listed in ``make_tf``): Conv1D = cross-correlation with SAME (extra pad on the right) /
CAUSAL / VALID padding and dilation, depthwise_conv2d SAME for the two-tap interpolation kernel (pad 0 left, 1 right),
conv1d_transpose with stride = filter width, l2_normalize with epsilon 1e-12, split, concat, tile, reshape, transpose, pad, tanh,
sigmoid, cumsum, floor-mod, gather.  The layer objects are created without running the Keras constructors (``object.__new__`` +
the attributes the ``call`` methods read, set from the model config exactly as custom_AE_layers.py:177-259 /
custom_pulsed_generator.py:459-493 would); ``TFPQMF`` is built by its real constructor.

What this pins: the *structure* the reference's code gives the computation -- pulse folding into channels, where the noise
channel goes, which half of a gate conv is tanh and which sigmoid, that the conditioning is added before the split, the
residual / skip split order and the last layer's skip-only case, `end` applied to the skip sum, the conditioning path
(sub-pixel conv unfold order, x10 interpolation weights and tail), the post 1x1, zero-stuffing / gain / padding of the PQMF
synthesis.  What it cannot pin is TensorFlow's arithmetic inside a primitive (summation order of a convolution): float outputs
are compared with a tolerance.

Output: tests/golden/reference_excitation.npz (committed); tests/test_reference_source.py checks the oracle against it,
tests/test_gpu_parity.py the CUDA path.
"""
import ast
import os
import sys
import types
import typing

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
F32 = np.float32


def _segments(path, names, class_name=None):
    """Source of the named top-level defs / classes (or methods of `class_name`), as written."""
    src = open(path).read()
    tree = ast.parse(src)
    bodies = tree.body
    if class_name is not None:
        bodies = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0].body
    out = {}
    for node in bodies:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            first = min([node.lineno] + [d.lineno for d in getattr(node, "decorator_list", [])])
            lines = src.splitlines()[first - 1:node.end_lineno]
            col = node.col_offset
            out[node.name] = "\n".join(ln[col:] if len(ln) >= col else ln for ln in lines)
    return out


# ---- NumPy float32 primitives ------------------------------------------------------------------------------------------
def conv1d(x, kernel, bias=None, padding="VALID", dilation=1, stride=1):
    """Keras Conv1D / tf.nn.conv1d: cross-correlation, channels-last, kernel (k, cin, cout)."""
    x, kernel = np.asarray(x, F32), np.asarray(kernel, F32)
    k = kernel.shape[0]
    span = dilation * (k - 1)
    pad = {"VALID": (0, 0), "SAME": (span // 2, span - span // 2), "CAUSAL": (span, 0)}[padding.upper()]
    xp = np.pad(x, ((0, 0), pad, (0, 0)))
    n_out = (xp.shape[1] - span - 1) // stride + 1
    out = np.zeros((x.shape[0], n_out, kernel.shape[2]), dtype=F32)
    for j in range(k):
        out += np.matmul(xp[:, j * dilation:j * dilation + (n_out - 1) * stride + 1:stride], kernel[j]).astype(F32)
    if bias is not None:
        out = out + np.asarray(bias, F32)
    return out


def depthwise_conv2d(x, filter, strides, padding, data_format="NHWC", dilations=None):
    assert padding == "SAME" and data_format == "NHWC" and x.shape[1] == 1 and filter.shape[0] == 1
    kw, c, mult = filter.shape[1], filter.shape[2], filter.shape[3]
    total = kw - 1
    xp = np.pad(x, ((0, 0), (0, 0), (total // 2, total - total // 2), (0, 0)))
    T = x.shape[2]
    out = np.zeros(x.shape[:3] + (c, mult), dtype=F32)
    for j in range(kw):
        out += xp[:, :, j:j + T, :, None] * filter[0, j][None, None, None]
    return out.reshape(x.shape[:3] + (c * mult,))               # output channel = c * multiplier + m


def conv1d_transpose(x, filters, output_shape, strides, padding="SAME"):
    w, cout, cin = filters.shape
    assert w == strides                                         # no overlap, no cropping
    B, T, _ = x.shape
    out = np.zeros((B, T * strides, cout), dtype=F32)
    for j in range(w):
        out[:, j::strides] = np.matmul(x, np.asarray(filters[j], F32).T)
    assert tuple(int(s) for s in output_shape) == out.shape
    return out


def l2_normalize(v, axis):
    sq = np.sum(np.square(v), axis=tuple(axis), keepdims=True, dtype=F32)
    return (v / np.sqrt(np.maximum(sq, F32(1e-12)))).astype(F32)


class Layer:
    """Stand-in for tf.keras.layers.Layer and the TF2C base layers: calling a layer calls its `call`."""
    dtype = np.float32

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class KerasConv1D:
    """The `conv1d_layer` member of TF2C_Conv1DWeightNorm: its kernel is assigned by the reference's call()."""

    def __init__(self, bias, padding="valid", dilation=1):
        self.kernel, self.bias, self.padding, self.dilation = None, np.asarray(bias, F32), padding, dilation

    def call(self, x):
        return conv1d(x, self.kernel, self.bias, self.padding, self.dilation)


def make_tf(state):
    def _f(x):
        return x if isinstance(x, np.ndarray) else np.asarray(x, dtype=F32)

    def gather(params, indices, axis=0, batch_dims=0):
        assert axis == 0                                        # array_ops.gather: the axis == 0 branch ignores batch_dims
        state.setdefault("gather", []).append(np.array(indices))
        return np.asarray(params)[indices]

    def split_kw(value, num_or_size_splits=None, axis=-1, *a):
        if isinstance(num_or_size_splits, (int, np.integer)):
            return list(np.split(value, num_or_size_splits, axis=axis))
        raise NotImplementedError

    def tf_split(value, num_or_size_splits, axis=0):
        return split_kw(value, num_or_size_splits, axis)

    def function(*a, **k):                                      # @tf.function(...) -> the undecorated function
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda fn: fn

    nn = types.SimpleNamespace(tanh=lambda x: np.tanh(_f(x)), sigmoid=lambda x: (F32(1) / (F32(1) + np.exp(-_f(x)))).astype(F32),
                               l2_normalize=l2_normalize, depthwise_conv2d=depthwise_conv2d, conv1d_transpose=conv1d_transpose,
                               conv1d=lambda x, f, stride=1, padding="VALID": conv1d(x, f, None, padding, 1, stride))
    keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=Layer))
    return types.SimpleNamespace(
        newaxis=None, float32=np.float32, int32=np.int32, int64=np.int64, bool=np.bool_, Tensor=np.ndarray,
        keras=keras, nn=nn, function=function, TensorSpec=lambda *a, **k: None, Variable=type("Variable", (), {}),
        random=types.SimpleNamespace(normal=lambda shape: state["noise"].reshape([int(s) for s in shape])),
        reshape=lambda x, shape: np.reshape(x, [int(s) for s in shape]),
        shape=lambda x: x.shape, transpose=lambda x, perm: np.transpose(x, perm),
        expand_dims=lambda x, axis: np.expand_dims(x, axis), tile=lambda x, reps: np.tile(x, reps),
        convert_to_tensor=lambda x, dtype=None: np.asarray(x).astype(dtype or F32),
        split=tf_split, gather=gather,
        cumsum=lambda x, axis=0: np.cumsum(x, axis=axis, dtype=x.dtype),
        pad=lambda x, paddings, **kw: np.pad(x, paddings),
        constant=lambda v, dtype=np.float32: np.asarray(v, dtype=dtype)[()],
        cast=lambda x, dtype: np.asarray(x).astype(dtype) if isinstance(x, np.ndarray) else dtype(x),
        range=lambda n: np.arange(n, dtype=np.int32), zeros=lambda shape, dtype=F32: np.zeros(shape, dtype=dtype),
        floor=lambda x: np.floor(_f(x)), abs=lambda x: np.abs(_f(x)), sin=lambda x: np.sin(_f(x)), sqrt=lambda x: np.sqrt(_f(x)),
        maximum=lambda a, b: np.maximum(_f(a), _f(b)), minimum=lambda a, b: np.minimum(_f(a), _f(b)),
        reduce_sum=lambda x, axis=None: np.sum(x, axis=axis, dtype=x.dtype),
        concat=lambda xs, axis=0: np.concatenate(list(xs), axis=axis),
        math=types.SimpleNamespace(log=lambda x: np.log(_f(x))))


def load_reference(state):
    tf = make_tf(state)
    import scipy.signal
    import scipy.signal.windows
    ss_shim = types.ModuleType("ss_shim")                        # SciPy >= 1.13 moved kaiser to scipy.signal.windows
    ss_shim.__dict__.update(scipy.signal.__dict__)
    ss_shim.kaiser = scipy.signal.windows.kaiser
    ns = {"tf": tf, "np": np, "layers": tf.keras.layers, "sys": sys, "copy": __import__("copy"),
          "TF2C_BasePretrainableLayer": Layer, "TF2C_BaseLayer": Layer, "activations": types.SimpleNamespace(get=lambda a: None),
          "ss": ss_shim}
    ns.update({k: getattr(typing, k) for k in ("Union", "Tuple", "List", "Dict", "Optional", "Any", "Sequence", "Callable")})
    model = os.path.join(REF, "MBExWN_NVoc/vocoder/model")
    lay = os.path.join(model, "tf2_components/layers")
    todo = [(os.path.join(lay, "conv_layers.py"), ["TF2C_Conv1DWeightNorm", "TF2C_Conv1DUpDownSample"], None),
            (os.path.join(lay, "support_layers.py"), ["_init_linear_interpolator_weights", "TF2C_LinInterpLayer"], None),
            (os.path.join(model, "tf_preprocess.py"), ["_design_prototype_filter", "TFPQMF"], None),
            (os.path.join(model, "tf_wavetable.py"), ["pad_axis"], None),
            (os.path.join(model, "tf_wavetable.py"), ["stable_cumsum_and_wrap", "call", "_linear_lookup"], "PulseWaveTable"),
            (os.path.join(model, "custom_pulsed_generator.py"), ["generate_excitation"], "MBExWN")]
    pulse_methods = {}
    for path, names, cls in todo:
        seg = _segments(path, set(names), cls)
        assert sorted(seg) == sorted(names), (path, sorted(seg))
        for name in names:                                      # definition order matters for the class hierarchy
            code = seg[name]
            if cls == "PulseWaveTable":
                code = code[code.index("def "):]                # @tf.function(input_signature=...) is not needed on a plain function
                local = dict(ns)
                exec(compile(code, f"{path}:{cls}.{name}", "exec"), local)
                pulse_methods[name] = local[name]
                continue
            exec(compile(code, f"{path}:{name}", "exec"), ns)
    ns["TF2C_Conv1DWeightNorm"].__module__ = "reference"
    # WaveNetAE / WaveNetAEBlock refer to the conv classes through these aliases inside their constructors only
    seg = _segments(os.path.join(model, "custom_AE_layers.py"), {"WaveNetAE", "WaveNetAEBlock"})
    for name in ("WaveNetAE", "WaveNetAEBlock"):
        exec(compile(seg[name], f"custom_AE_layers.py:{name}", "exec"), ns)
    ns["PulseWaveTable"] = type("PulseWaveTable", (Layer,), pulse_methods)
    return tf, ns


def build_generator(ns, tf, plan, weights, tables_grid):
    """The object graph generate_excitation walks, with the attributes MBExWN.__init__ would have set."""
    wn = plan.wavenet
    padding = "CAUSAL" if wn.causal else "SAME"

    def conv(cls, name, padding="valid", dilation=1, **extra):
        o = object.__new__(cls)
        o.use_equalized_lr, o.use_weight_norm, o.kernel_norm_axes = False, True, [0, 1]      # conv_layers.py:77-79
        o.pretrain_activations, o.activation = False, None
        o.v, o.g = np.asarray(weights[f"{name}/v"], F32), np.asarray(weights[f"{name}/g"], F32)
        o.conv1d_layer = KerasConv1D(weights[f"{name}/bias"], padding, dilation)
        for k, v in extra.items():
            setattr(o, k, v)
        return o

    CW, CU = ns["TF2C_Conv1DWeightNorm"], ns["TF2C_Conv1DUpDownSample"]
    blocks = []
    for b in plan.blocks:
        base = b.name + "_WNBlock_WN"
        w = object.__new__(ns["WaveNetAE"])
        w.n_layers, w.n_ch_groups, w.activation = b.n_layers, 1, {0: "gtu", 1: "glu", 2: "gfu", 3: "gsu"}[b.gate]
        w.pre_cond_layers, w.cond_conv_upsampling = [], b.cond_conv_up
        w.return_activations, w.return_activations_mask = None, [False] * b.n_layers          # custom_AE_layers.py:262-265
        w.start, w.end = conv(CW, f"{base}/start"), conv(CW, f"{base}/end")
        w.cond_layer = conv(CU, f"{base}/cond_", padding, up_sample=True, down_sample=False, factor=b.cond_conv_up)
        li = object.__new__(ns["TF2C_LinInterpLayer"])                                       # :223-225
        li.upsampling_factor, li.num_pad_end, li.drop_last, li.last_size, li.single_channel_mode = b.cond_lin_up, 1, True, 0, False
        li.kernel = ns["_init_linear_interpolator_weights"]([1, 2, 2 * b.c, b.cond_lin_up], np.float32)
        w.cond_lin_upsampling_layer = li
        w.conv_layers = [conv(CW, f"{base}/conv1D_{i}", padding, d) for i, d in enumerate(b.dilations)]
        w.res_skip_layers = [conv(CW, f"{base}/res_skip_{i}") for i in range(b.n_layers)]
        blk = object.__new__(ns["WaveNetAEBlock"])
        blk.wavenet = w
        blk.up_down_sample = conv(CU, b.up_name, padding, up_sample=True, down_sample=False, factor=b.up) if b.up > 1 else None
        blocks.append(blk)

    tables, grid, nominal = tables_grid
    pg = object.__new__(ns["PulseWaveTable"])
    pg.sample_rate, pg.use_sinusoid_as_fun, pg.add_subharm_chans = plan.pulse_rate, False, plan.subharm
    pg.pulse_sync_gain_avg, pg.no_interp, pg.wavetables, pg.n_period, pg.nominalF0 = False, False, tables, int(tables.shape[0] - 1), nominal
    pg.minTranspositionFactorInGrid = tf.constant(np.min(grid) / nominal, tf.float32)       # tf_wavetable.py:283-284, :305
    pg.maxTranspositionFactorInGrid = tf.constant(np.max(grid) / nominal, tf.float32)
    pg.grid_f0_diff_norm_factor = 1. / tf.math.log(1.25)

    pulse_pqmf = None
    if plan.pulse_pqmf_cfg is not None:                                                    # custom_pulsed_generator.py:499-501
        pulse_pqmf = ns["TFPQMF"](**plan.pulse_pqmf_cfg, do_synthesis=False, name="PC_PQMFilterBank")
    gen = types.SimpleNamespace(pulse_generator=pg, pulse_pqmf=pulse_pqmf, pulse_channels=plan.pulse_channels,
                                pp_mod_subnet_noise_channel_sigma=plan.noise_sigma, pp_waveNetBlocks=blocks,
                                wn_post_net=[conv(CW, plan.post_name)],
                                pqmf=ns["TFPQMF"](**plan.pqmf_cfg, do_synthesis=True, name="PQMFilterBank"))   # real constructor
    return gen


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import synthetic_mel, synthetic_noise

    state = {}
    tf, ns = load_reference(state)
    gold = np.load(os.path.join(HERE, "reference_init_dsp.npz"))
    out = {}
    cases = {"speech": ({}, 3), "blocks_2x1": ({"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                                                "pp_mod_subnet_channel_factors": [0.5, 0.25]}, 5)}
    for tag, (extra, seed) in cases.items():
        hp = read_config(get_config_file("SPEECH"))
        hp["mbexwn_config"].update(extra)
        plan = build_plan(hp)
        weights = W.init_synthetic(plan, seed=seed)
        assert np.array_equal(plan.wavetables.tables, gold["wt_sp_tables"])        # the bank the reference's own code builds
        gen = build_generator(ns, tf, plan, weights, (gold["wt_sp_tables"], gold["wt_sp_grid"], gold["wt_sp_cfg"][3]))
        T = 12
        mel = np.stack([synthetic_mel(T, i) for i in range(2)]).astype(F32)
        noise = np.stack([synthetic_noise(T * plan.steps_per_frame, i) for i in range(2)]).astype(F32)
        t = np.arange(T * plan.pulse_per_frame) / (T * plan.pulse_per_frame)
        f0 = np.stack([90.0 * (4.0 ** t), 310.0 - 200.0 * t]).astype(F32)
        state["noise"] = noise
        state["gather"] = []
        exc = ns["generate_excitation"](gen, mel, f0)
        assert exc.shape == (2, T * plan.hop) and exc.dtype == np.float32, (exc.shape, exc.dtype)
        out[f"{tag}_seed"] = np.array(seed)
        out[f"{tag}_mel"], out[f"{tag}_noise"], out[f"{tag}_f0"], out[f"{tag}_excitation"] = mel, noise, f0, exc
        out[f"{tag}_index"] = state["gather"][0][:, :, 0].astype(np.int32)
        # the WaveNet blocks alone, from the same folded pulse / noise rows (what the blocks loop of generate_excitation sees)
        x = gen.pulse_generator(f0).reshape(2, -1, plan.pulse_channels)
        x = np.concatenate((x, plan.noise_sigma * noise.reshape(2, -1, 1)), axis=-1).astype(F32)
        for bl in gen.pp_waveNetBlocks:
            x = bl((x, mel))
        out[f"{tag}_wn_out"] = x
        out[f"{tag}_subbands"] = gen.wn_post_net[0](x)
        print(tag, "excitation", exc.shape, "peak", float(np.abs(exc).max()), "blocks out", x.shape)
    path = os.path.join(HERE, "reference_excitation.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
