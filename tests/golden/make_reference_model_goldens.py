#!/usr/bin/env python
"""Golden vectors from the reference's model OBJECT built by its own constructors (runs only where /root/reference is mounted).

The last step of the series make_reference_{pulse,excitation,forward}_goldens.py: here nothing of the model is assembled by hand.
The classes ``MBExWN`` (custom_pulsed_generator.py:151-925, constructor included: rate algebra, sub-net construction, WaveNet blocks
with their conditioning factors and dilation schedule, lifter bank, F0 smoothing kernel), ``WaveNetAEBlock`` / ``WaveNetAE``
(constructors included), ``PulseWaveTable`` (constructor included: the wavetable bank is built by it), ``TFPQMF``, ``TFPad1d``,
``TF2C_LinInterpLayer``, ``ActivationLayer`` and ``PaNWaveNet.infer`` (wavegen_1d.py:483-526) are compiled unmodified from
/root/reference and the model is created as ``MBExWN(preprocess_config=..., **mbexwn_config)`` from this package's config.yaml (the
reference's own keyword arguments).  Stand-ins: the TensorFlow primitives (NumPy float32, see the sibling scripts), the Keras base
``Layer`` (name, lazy build, add_weight), Keras' PReLU, and the constructor of the two weight-normalised conv classes, which fetches
``v`` / ``g`` / ``bias`` by layer name instead of creating variables (their ``call`` is the reference's).

Output: tests/golden/reference_model.npz (committed); tests/test_reference_source.py checks the oracle against it.
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
CASES = {"speech": ({}, 0, 11),
         "blocks_2x1": ({"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                         "pp_mod_subnet_channel_factors": [0.5, 0.25]}, 5, 9),
         "lifter_causal": ({"ps_env_order_scale": 1.5, "force_causal": True}, 6, 9)}


def load_model_classes(state, weights):
    tf, ns = FW.load_forward(state, weights)
    sys.path.insert(0, X.REF)
    from MBExWN_NVoc.glottis.FglotspecLF import FglotspecLF                        # plain NumPy, imported as is
    import scipy.signal
    import scipy.signal.windows
    np_shim = types.ModuleType("np_shim")                                          # NumPy 2 removed these aliases
    np_shim.__dict__.update(np.__dict__)
    np_shim.int, np_shim.float = int, float
    np_shim.cast = {np.int32: lambda x: np.asarray(x).astype(np.int32)}
    ss_shim = types.ModuleType("ss_shim")
    ss_shim.__dict__.update(scipy.signal.__dict__)
    ss_shim.kaiser = scipy.signal.windows.kaiser
    model = os.path.join(X.REF, "MBExWN_NVoc/vocoder/model")
    # tf_wavetable.py: module-level helpers + the whole PulseWaveTable class, in their own namespace (NumPy / SciPy shims)
    wt_ns = dict(ns, np=np_shim, ss=ss_shim, FglotspecLF=FglotspecLF, print=lambda *a, **k: None)
    wt_path = os.path.join(model, "tf_wavetable.py")
    names = ["get_pulse_lowpass_kaiser", "get_min_phase_spectrum", "get_LFpulse", "pad_axis", "PulseWaveTable"]
    seg = X._segments(wt_path, set(names))
    assert sorted(seg) == sorted(names)
    for name in names:
        exec(compile(seg[name], f"{wt_path}:{name}", "exec"), wt_ns)
    ns["PulseWaveTable"] = wt_ns["PulseWaveTable"]
    # WaveNetAE(.Block) again, now on the base class that records names (their constructors run this time)
    ae_path = os.path.join(model, "custom_AE_layers.py")
    seg = X._segments(ae_path, {"WaveNetAE", "WaveNetAEBlock"})
    for name in ("WaveNetAE", "WaveNetAEBlock"):
        exec(compile(seg[name], f"{ae_path}:{name}", "exec"), ns)
    ns["log_to_db"] = 20 * np.log10(np.exp(1))                                     # custom_pulsed_generator.py:26
    ns["ParamSchedule"] = None
    gen_path = os.path.join(model, "custom_pulsed_generator.py")
    exec(compile(X._segments(gen_path, {"MBExWN"})["MBExWN"], gen_path + ":MBExWN", "exec"), ns)
    wg_path = os.path.join(model, "wavegen_1d.py")
    local = dict(ns)
    exec(compile(X._segments(wg_path, {"infer"}, "PaNWaveNet")["infer"], wg_path + ":PaNWaveNet.infer", "exec"), local)
    ns["infer"] = local["infer"]
    return tf, ns


def main():
    if not os.path.isdir(X.REF):
        print("reference not mounted; nothing to do")
        return 1
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import synthetic_mel, synthetic_noise

    out = {}
    for tag, (extra, seed, T) in CASES.items():
        hp = read_config(get_config_file("SPEECH"))
        hp["mbexwn_config"].update(extra)
        plan = build_plan(hp)
        weights = W.init_synthetic(plan, seed=seed)
        state = {}
        tf, ns = load_model_classes(state, weights)
        mc = dict(hp["mbexwn_config"])
        model = ns["MBExWN"](preprocess_config=hp["preprocess_config"], use_tf25_compatible_implementation=True, quiet=True, **mc)
        # what the constructor derived, against this package's plan
        assert model.spect_to_pulse_upsampling_factor == plan.pulse_per_frame and model.fft_size == plan.fft_size
        assert model.stft_win_size == plan.stft_win and model.n_period if hasattr(model, "n_period") else True
        assert np.array_equal(np.asarray(model.pulse_generator.wavetables), plan.wavetables.tables), "wavetable bank"
        for blk, spec in zip(model.pp_waveNetBlocks, plan.blocks):
            wn = blk.wavenet
            assert (wn.n_channels, wn.cond_conv_upsampling, wn.cond_lin_upsampling) == (spec.c, spec.cond_conv_up, spec.cond_lin_up)
            assert [l.conv1d_layer.dilation for l in wn.conv_layers] == spec.dilations
        # attribute tree of the reference object = the object-graph paths of a Keras checkpoint of it (tf_checkpoint.py writes /
        # reads those paths): position, kind and name of every member of the two sub-net layer lists
        table = []
        for attr in ("pp_subnet_layers", "ps_subnet_layers"):
            for i, layer in enumerate(getattr(model, attr)):
                kind = "conv" if hasattr(layer, "v") else "prelu" if hasattr(layer, "alpha") else type(layer).__name__
                key = getattr(layer, "weight_key", None) or getattr(layer, "name", "")
                table.append(f"{attr}/{i}|{kind}|{key}")
        for ib, blk in enumerate(model.pp_waveNetBlocks):
            for attr in ("start", "end", "cond_layer", "conv_layers", "res_skip_layers"):
                assert hasattr(blk.wavenet, attr), attr
            table.append(f"pp_waveNetBlocks/{ib}|block|{'up_down_sample' if blk.up_down_sample is not None else ''}")
        assert len(model.wn_post_net) == 1
        out[f"{tag}_layer_table"] = np.array(table)
        shell = types.SimpleNamespace(segment_length=0, spect_hop_size=plan.hop, norm_mel_components=None, block=model)
        mel = np.stack([synthetic_mel(T, 80 + i) for i in range(2)]).astype(F32)
        noise = np.stack([synthetic_noise(T * plan.steps_per_frame, 80 + i) for i in range(2)]).astype(F32)
        state["noise"], state["gather"] = noise, []
        signal, pp = ns["infer"](shell, mel, synth_length=T * plan.hop, return_F0=True)
        pp = dict((k, v) for k, v in pp)
        assert signal.shape == (2, T * plan.hop) and signal.dtype == np.float32, (signal.shape, signal.dtype)
        out[f"{tag}_seed"], out[f"{tag}_mel"], out[f"{tag}_noise"] = np.array(seed), mel, noise
        out[f"{tag}_F0"], out[f"{tag}_waveform"] = model.generate_f0(mel), signal
        out[f"{tag}_index"] = state["gather"][0][:, :, 0].astype(np.int32)
        if plan.env_order_scale:
            out[f"{tag}_lifters"] = np.asarray(model.ps_cepstral_windows, F32)
            out[f"{tag}_lifter_grid"] = np.asarray(model.ps_cepstral_windows_log10f0, F32)
        print(tag, "waveform", signal.shape, "peak", float(np.abs(signal).max()), "PP", sorted(pp))
    path = os.path.join(HERE, "reference_model.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
