#!/usr/bin/env python
"""Golden vectors of the WHOLE inference branch -- mel in, waveform out -- from the REAL reference source (runs only where
/root/reference is mounted).

Extends tests/golden/make_reference_excitation_goldens.py (same method: the reference's source compiled unmodified from
/root/reference, executed over NumPy float32 stand-ins for the TensorFlow primitives) by

* ``generate_subnet_from_specs``                 custom_pulsed_generator.py:38-148   -- the spec grammar of the F0 and VTF sub-nets is
  *parsed by the reference's code*; the layer classes it instantiates are the reference's own ``TFPad1d`` (custom_layers.py:20-71),
  ``TF2C_LinInterpLayer`` and ``ActivationLayer`` (custom_AE_layers.py:21-109), built by their real constructors, and sub-classes of
  the reference's ``TF2C_Conv1DWeightNorm`` / ``TF2C_Conv1DUpDownSample`` whose constructor only looks the weights up by layer name (the
  reference's ``call`` does the work); Keras' PReLU(shared_axes=[1]) is a stand-in (max(x, 0) + alpha min(x, 0));
* ``MBExWN.call`` (inference branch), ``generate_f0``, ``generate_specenv``, ``_get_cepstral_windows``   custom_pulsed_generator.py:556-855,
  :507-525.

Additional stand-ins, written from the TensorFlow documentation: ``tf.signal.stft`` (frames of frame_length every frame_step,
periodic Hann window, zero-padded rfft, pad_end False), ``tf.signal.inverse_stft`` (irfft, truncate to frame_length, window,
overlap-add), ``tf.signal.inverse_stft_window_fn`` (forward window / sum of its squares over the overlapping frames),
``tf.signal.rfft``, ``tf.complex`` / ``exp`` / ``real`` / ``imag`` / ``tanh``, ``tf.round`` (half to even), ``tf.pad`` with modes.
The lifter bank and the F0 smoothing kernel (numpy inside MBExWN.__init__, :404-450) come from mbexwn_vocoder_b200.dsp_init.

Output: tests/golden/reference_forward.npz (committed); tests/test_reference_source.py checks the oracle against it,
tests/test_gpu_parity.py the CUDA path.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_excitation_goldens as X                                      # noqa: E402

F32, C64 = np.float32, np.complex64
REF = X.REF


# ---- tf.signal, from the documented behaviour --------------------------------------------------------------------------------
def hann_window(n, periodic=True, dtype=F32):
    assert periodic
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(F32)


def stft(signals, frame_length, frame_step, fft_length=None, window_fn=hann_window, pad_end=False, name=None):
    assert not pad_end
    n = 1 + (signals.shape[-1] - frame_length) // frame_step
    idx = np.arange(frame_length)[None, :] + frame_step * np.arange(n)[:, None]
    frames = signals[..., idx] * window_fn(frame_length)
    return np.fft.rfft(frames.astype(F32), n=fft_length, axis=-1).astype(C64)


def overlap_and_add(frames, step):
    n, length = frames.shape[-2], frames.shape[-1]
    out = np.zeros(frames.shape[:-2] + ((n - 1) * step + length,), dtype=F32)
    for i in range(n):
        out[..., i * step:i * step + length] += frames[..., i, :]
    return out


def inverse_stft(stfts, frame_length, frame_step, fft_length=None, window_fn=hann_window, name=None):
    real = np.fft.irfft(stfts, n=fft_length, axis=-1).astype(F32)[..., :frame_length]
    return overlap_and_add(real * window_fn(frame_length, dtype=F32), frame_step)


def inverse_stft_window_fn(frame_step, forward_window_fn=hann_window, name=None):
    def fn(frame_length, dtype=F32):
        fw = forward_window_fn(frame_length)
        denom = np.square(fw)
        overlaps = -(-frame_length // frame_step)
        denom = np.pad(denom, (0, overlaps * frame_step - frame_length)).reshape(overlaps, frame_step)
        denom = np.tile(denom.sum(0, keepdims=True), (overlaps, 1)).reshape(-1)
        return (fw / denom[:frame_length]).astype(F32)
    return fn


def extend_tf(tf, weights):
    pad0 = tf.pad

    def pad(x, paddings, mode="CONSTANT", **kw):
        paddings = [tuple(int(v) for v in p) for p in paddings]
        return np.pad(x, paddings, mode={"CONSTANT": "constant", "SYMMETRIC": "symmetric", "REFLECT": "reflect"}[mode.upper()])

    class PReLU(X.Layer):                                      # tf.keras.layers.PReLU(shared_axes=[1]): one slope per channel
        def __init__(self, alpha_initializer=None, shared_axes=None, name=None, **kw):
            assert list(shared_axes) == [1]
            self.weight_key = name
            self.alpha = np.asarray(weights[f"{name}/alpha"], F32)

        def call(self, x):
            return (np.maximum(x, F32(0)) + self.alpha * np.minimum(x, F32(0))).astype(F32)

    del pad0
    tf.pad = pad
    tf.complex64 = C64
    tf.complex = lambda re, im: (np.asarray(re, F32) + 1j * np.asarray(im, F32)).astype(C64)
    tf.exp = lambda x: np.exp(x).astype(x.dtype)
    tf.round = lambda x: np.round(x)                           # half to even, like tf.round
    tf.assert_equal = lambda a, b, msg=None: None if np.all(a == b) else (_ for _ in ()).throw(AssertionError(msg))
    tf.stop_gradient = lambda x: x
    tf.reduce_mean = lambda x, axis=None, keepdims=False: np.mean(x, axis=axis, keepdims=keepdims, dtype=F32)
    tf.ones = lambda shape, dtype=F32: np.ones(shape, dtype=dtype)
    tf.math.tanh = lambda x: np.tanh(x).astype(F32)
    tf.math.real = lambda x: np.real(x).astype(F32)
    tf.math.imag = lambda x: np.imag(x).astype(F32)
    tf.signal = types.SimpleNamespace(stft=stft, inverse_stft=inverse_stft, inverse_stft_window_fn=inverse_stft_window_fn,
                                      hann_window=hann_window, rfft=lambda x: np.fft.rfft(x, axis=-1).astype(C64))
    tf.nn.conv1d = lambda x, f, stride=1, padding="VALID", dilations=1: X.conv1d(x, f, None, padding, 1, stride)
    tf.keras.layers.PReLU = PReLU
    tf.keras.initializers = types.SimpleNamespace(RandomNormal=lambda **k: None, Constant=lambda v: None, constant=lambda v: None)
    return tf


SCOPE = [None]          # name of the WaveNetAE under construction (Keras name scope of its "start", "conv1D_<i>" ... sub-layers)


def load_forward(state, weights):
    tf, ns = X.load_reference(state)
    extend_tf(tf, weights)
    model = os.path.join(REF, "MBExWN_NVoc/vocoder/model")
    # real classes with their real constructors: the base-class stand-in builds lazily on the first call, as Keras does
    class Layer(X.Layer):
        def __init__(self, *a, trainable=True, name=None, **k):
            self.trainable, self.name, self._built = trainable, name, False
            if type(self).__name__ == "WaveNetAE":
                SCOPE[0] = name

        def add_weight(self, name=None, shape=None, initializer=None, dtype=None, trainable=False):
            return initializer(shape, dtype or F32)

        def __call__(self, x, *a, **k):
            # leaf layers build on their first call, as in Keras; composite models (MBExWN, WaveNetAE ...) only pass the call on
            if (not getattr(self, "_built", True) and hasattr(self, "build_or_compute_output_shape")
                    and type(self).__name__ in ("TF2C_LinInterpLayer", "TFPad1d", "ActivationLayer")):
                self.build_or_compute_output_shape(tuple(x.shape), do_build=True)
                self._built = True
            return self.call(x, *a, **k)

    ns["TF2C_BaseLayer"] = ns["TF2C_BasePretrainableLayer"] = Layer
    ns["layers"] = types.SimpleNamespace(Layer=Layer, PReLU=tf.keras.layers.PReLU)
    for path, names in ((os.path.join(model, "tf2_components/layers/support_layers.py"), ["TF2C_LinInterpLayer"]),
                        (os.path.join(model, "custom_layers.py"), ["TFPad1d"]),
                        (os.path.join(model, "custom_AE_layers.py"), ["ActivationLayer"]),
                        (os.path.join(model, "custom_pulsed_generator.py"), ["get_missing_upsampling_factor", "generate_subnet_from_specs"])):
        seg = X._segments(path, set(names))
        assert sorted(seg) == sorted(names), (path, sorted(seg))
        if "TFPad1d" in names:
            ns.setdefault("TF2C_BaseLayer", Layer)
        for name in names:
            exec(compile(seg[name], f"{path}:{name}", "exec"), ns)
    ns["LinInterpLayer"] = ns["TF2C_LinInterpLayer"]

    # sub-net convs: the reference's call() on top of a constructor that only fetches the weights by layer name
    def named(cls):
        class Named(cls):
            def __init__(self, filters, kernel_size=1, padding="valid", use_weight_norm=True, name=None, factor=1, up_sample=None,
                         **kw):
                assert use_weight_norm
                self.use_equalized_lr, self.use_weight_norm, self.kernel_norm_axes = False, True, [0, 1]
                self.pretrain_activations, self.activation = False, None
                if f"{name}/v" not in weights:                 # a sub-layer of a WaveNetAE built by the reference's constructor
                    name = f"{SCOPE[0]}/{name}"
                self.weight_key = name
                self.v, self.g = np.asarray(weights[f"{name}/v"], F32), np.asarray(weights[f"{name}/g"], F32)
                assert self.v.shape[0] == kernel_size and self.v.shape[2] == filters * (factor if up_sample else 1), name
                self.conv1d_layer = X.KerasConv1D(weights[f"{name}/bias"], padding, kw.get("dilation_rate", 1))
                self.up_sample, self.down_sample, self.factor = up_sample, False, factor

            def __call__(self, x):
                return self.call(x)
        return Named

    ns["TF2C_Conv1DWeightNorm"], ns["TF2C_Conv1DUpDownSample"] = named(ns["TF2C_Conv1DWeightNorm"]), named(ns["TF2C_Conv1DUpDownSample"])
    methods = X._segments(os.path.join(model, "custom_pulsed_generator.py"),
                          {"call", "generate_f0", "generate_specenv", "_get_cepstral_windows", "generate_multiband_gain"}, "MBExWN")
    assert len(methods) == 5
    for name, code in methods.items():
        local = dict(ns)
        exec(compile(code, f"custom_pulsed_generator.py:MBExWN.{name}", "exec"), local)
        ns["MBExWN_" + name] = local[name]
    return tf, ns


def build_model(ns, tf, hp, plan, weights, gen):
    """`self` of MBExWN.call: the excitation object graph of the other script + what MBExWN.__init__ sets for the F0 / VTF side."""
    mc = hp["mbexwn_config"]
    act_kwargs = {"alpha_initializer": None, "shared_axes": [1]}                           # custom_pulsed_generator.py:247-250
    common = dict(activation=tf.keras.layers.PReLU, force_causal=bool(mc.get("force_causal", False)),
                  remove_inactive_pad_layers=bool(mc.get("remove_inactive_pad_layers", False)),
                  use_tf25_compatible_implementation=True, **act_kwargs)
    m = gen
    m.pp_subnet_layers, _ = ns["generate_subnet_from_specs"](                              # :292-305
        mc["pp_subnet"], base_name="PulsPar", final_nks=1, final_n_channels=1,
        final_activation=mc.get("pp_activation", "soft_sigmoid"), target_ups=plan.pulse_per_frame,
        pad_to_valid=bool(mc.get("pp_subnet_use_valid_padding", False)), **common)
    ps_use_stft = bool(mc.get("ps_use_stft", True))
    m.ps_subnet_layers, _ = ns["generate_subnet_from_specs"](                              # :412-426
        mc["ps_subnet"], base_name="PS", final_nks=1, final_n_channels=plan.n_ceps if ps_use_stft else plan.subbands,
        final_activation=None,
        pad_to_valid=bool(mc.get("ps_subnet_use_valid_padding", False)), weight_init_scale=0.01, **common)
    m.pp_max_frequency, m.pp_min_frequency = plan.f0_max, plan.f0_min
    m.spect_to_pulse_upsampling_factor, m.spect_hop_size = plan.pulse_per_frame, plan.hop
    m.sample_rate, m.pulse_rate = plan.sample_rate, plan.pulse_rate
    m.ps_use_stft, m.ps_off, m.dump_controls = ps_use_stft, False, False
    if not ps_use_stft:                                                                    # :452-453
        m.ps_gain_interpolator = ns["LinInterpLayer"](upsampling_factor=plan.hop, num_pad_end=1)
    m.stft_win_size, m.fft_size, m.stft_win = plan.stft_win, plan.fft_size, tf.signal.hann_window
    m.ps_env_order_scale, m.psns_use_cepstral_loss_constraint = plan.env_order_scale, False
    if plan.env_order_scale:
        m.ps_cepstral_windows_log10f0, m.ps_cepstral_windows = np.asarray(plan.lifter_log10f0, F32), np.asarray(plan.lifters, F32)
    m.frequency_smoothing_kernel = np.asarray(plan.f0_smooth, F32)[:, None, None]          # :404-406
    m.log_to_log10 = 1 / np.log(10)                                                        # :503
    m.spect_filters_preserve_energy, m.psns_gain_loss_weight = bool(mc.get("spect_filters_preserve_energy", False)), 0
    m.filter_max_log_range = plan.filter_max_log_range
    m.pulse_noise_floor_mag, m.stft_coh_loss_weight = None, 0
    m.pp_subnet_training_only, m.pp_teacher_forcing_schedule, m.pulse_rate_factor = False, None, plan.pulse_rate_factor
    for name in ("generate_f0", "generate_specenv", "_get_cepstral_windows", "generate_excitation", "generate_multiband_gain"):
        fn = ns.get("MBExWN_" + name) or ns[name]
        setattr(m, name, types.MethodType(fn, m))
    return m


def main():
    if not os.path.isdir(REF):
        print("reference not mounted; nothing to do")
        return 1
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import synthetic_mel, synthetic_noise

    gold = np.load(os.path.join(HERE, "reference_init_dsp.npz"))
    out = {}
    # the scheme model, and the same with the F0-dependent cepstral lifter switched on (ps_env_order_scale, :434-450, :800-812)
    pulse_pqmf = {"pulse_channels_use_pqmf": True,
                  "pulse_channels_multi_band_config": {"subbands": 5, "taps": 40, "cutoff_ratio": 0.11, "beta": 8.0}}
    cases = (("speech", "SPEECH", 0, 21, {}), ("speech_lifter", "SPEECH", 1, 17, {"ps_env_order_scale": 2.0}),
             # variants of the path, short utterances
             ("band_gain_centered", "SPEECH", 2, 9, {"ps_use_stft": False, "spect_filters_preserve_energy": True}),
             ("causal", "SPEECH", 3, 9, {"force_causal": True}),
             ("pulse_pqmf_subharm", "SPEECH", 4, 9,
              dict(pulse_pqmf, wavetable_config={"nominalF0": 60, "maxF0": 550, "add_subharm_chans": 1})))
    for tag, model_id, seed, T, extra in cases:
        hp = read_config(get_config_file(model_id))
        hp["mbexwn_config"].update(extra)
        plan = build_plan(hp)
        weights = W.init_synthetic(plan, seed=seed)
        state = {}
        tf, ns = load_forward(state, weights)
        gen = X.build_generator(ns, tf, plan, weights, (gold["wt_sp_tables"], gold["wt_sp_grid"], gold["wt_sp_cfg"][3]))
        model = build_model(ns, tf, hp, plan, weights, gen)
        mel = np.stack([synthetic_mel(T, 40 + i) for i in range(2)]).astype(F32)
        noise = np.stack([synthetic_noise(T * plan.steps_per_frame, 40 + i) for i in range(2)]).astype(F32)
        state["noise"], state["gather"] = noise, []
        (signal,), pp = ns["MBExWN_call"](model, mel, return_PP=True)
        pp = dict((k, v) for k, v in pp)
        f0 = model.generate_f0(mel)
        assert signal.shape == (2, T * plan.hop) and signal.dtype == np.float32, (signal.shape, signal.dtype)
        out[f"{tag}_seed"], out[f"{tag}_mel"], out[f"{tag}_noise"] = np.array(seed), mel, noise
        out[f"{tag}_F0"], out[f"{tag}_waveform"] = f0, signal
        if "PSig" in pp:
            out[f"{tag}_excitation"] = pp["PSig"]
        if T > 10:                                             # the |VTF| tap only for the two main cases (size)
            out[f"{tag}_vtf_mag"] = pp["PS"].astype(F32)
        out[f"{tag}_index"] = state["gather"][0][:, :, 0].astype(np.int32)
        if plan.env_order_scale:
            out[f"{tag}_lifter_index"] = state["gather"][1].astype(np.int32)
        print(tag, "waveform", signal.shape, "peak", float(np.abs(signal).max()), "F0", float(f0.min()), float(f0.max()),
              "layers", len(model.pp_subnet_layers), len(model.ps_subnet_layers))
    path = os.path.join(HERE, "reference_forward.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
