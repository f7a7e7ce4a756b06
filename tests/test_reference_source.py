"""The oracle and the host logic against goldens produced by the REAL reference source.  tests/golden/reference_*.npz hold what the
reference's own code -- compiled unmodified from /root/reference by tests/golden/make_reference_*_goldens.py and executed over NumPy
stand-ins for the TensorFlow primitives -- returns for the pulse generator, the excitation branch, the whole inference branch, the
model object built by its own constructors, NormMelComponents and MELInverter.scale_mel.  Integer indices and the wrapped phase must
match bit for bit, float outputs to float32 rounding.  The CUDA path is held to the same files in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

from mbexwn_vocoder_b200 import get_config_file, weights as W
from mbexwn_vocoder_b200.config import read_config
from mbexwn_vocoder_b200.plan import build_plan
from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_pulse.npz"))
MODEL = {"sp": "SPEECH", "vo": "SING"}                  # wavetable_config (nominalF0 60, maxF0 550 / 1400), pulse rate 8 kHz


def _oracle(tag, subharm):
    hp = read_config(get_config_file(MODEL[tag]))
    if subharm:
        hp["mbexwn_config"]["wavetable_config"] = dict(hp["mbexwn_config"]["wavetable_config"], add_subharm_chans=subharm)
    plan = build_plan(hp)
    return plan, OracleMBExWN(hp, W.init_synthetic(plan, seed=0), torch.float32)


@pytest.mark.parametrize("tag", ["sp", "vo"])
def test_oracle_pulse_generator_matches_the_reference_source(tag):
    plan, orc = _oracle(tag, 0)
    assert np.array_equal(orc.wt.tables, np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_init_dsp.npz"))[f"wt_{tag}_tables"])
    for n in (2300, 1000, 700):
        key = f"{tag}_s0_{n}"
        f0 = GOLD[key + "_f0"]
        got = orc.pulse_generator(f0)
        assert got["phase"].dtype == np.float32
        assert np.array_equal(got["phase"], GOLD[key + "_phase"]), key           # stable_cumsum_and_wrap, bit for bit
        assert np.array_equal(got["index"], GOLD[key + "_index"]), key           # floor(phase * n_period) as int32
        ref = GOLD[key + "_audio"][:, :, 0]
        assert np.abs(got["pulse"] - ref).max() <= 1e-6 * np.abs(ref).max(), key


def test_oracle_subharmonic_channels_match_the_reference_source():
    """add_subharm_chans = 2 (tf_wavetable.py:554-559) through generate_excitation: WaveNet input rows are the reference's
    (B, T, 3) output folded by pulse_channels (custom_pulsed_generator.py:893)."""
    plan, orc = _oracle("sp", 2)
    key = "sp_s2_700"
    f0 = GOLD[key + "_f0"]
    T = 700 // plan.pulse_per_frame
    mel = torch.as_tensor(np.stack([synthetic_mel(T, i) for i in range(3)]))
    nz = torch.as_tensor(np.stack([synthetic_noise(T * plan.steps_per_frame, i) for i in range(3)])).reshape(3, -1, 1)
    taps = {}
    orc.generate_excitation(mel, torch.as_tensor(f0), nz, taps)
    x = taps["wn_in"].numpy()[:, :, :-1].reshape(3, 700, 3)
    ref = GOLD[key + "_audio"]
    assert np.array_equal(taps["index"], GOLD[key + "_index"])
    assert np.abs(x - ref).max() <= 2e-6


def test_goldens_cover_chunk_boundaries_and_the_table_range():
    """The vectors exercise what they are meant to pin: more than one cumsum chunk, a ragged last chunk, phase wraps, every
    table of the bank selected by the sweep, and indices up to the last table row."""
    idx = GOLD["vo_s0_2300_index"]
    ph = GOLD["vo_s0_2300_phase"]
    assert idx.min() == 0 and idx.max() == 511 and ph.min() >= 0 and ph.max() < 1
    assert (np.diff(ph[1]) < 0).sum() > 100                                      # the sweep wraps many times
    f0 = GOLD["vo_s0_2300_f0"][1]
    assert f0.min() < 50 and f0.max() > 1350


# ---- the whole excitation branch (pulse -> WaveNet blocks -> post net -> PQMF synthesis) --------------------------------------
EXC = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_excitation.npz"))
EXC_CASES = {"speech": {}, "blocks_2x1": {"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                                          "pp_mod_subnet_channel_factors": [0.5, 0.25]}}


def excitation_case(tag):
    """(hparams, plan, weights) of a case of tests/golden/make_reference_excitation_goldens.py."""
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(EXC_CASES[tag])
    plan = build_plan(hp)
    return hp, plan, W.init_synthetic(plan, seed=int(EXC[f"{tag}_seed"]))


@pytest.mark.parametrize("tag", sorted(EXC_CASES))
def test_oracle_excitation_branch_matches_the_reference_source(tag):
    """tests/golden/reference_excitation.npz = MBExWN.generate_excitation, WaveNetAE(.Block).call, the weight-norm / sub-pixel conv
    and LinInterp call methods, TFPQMF (constructor + synthesis) and the pulse generator, executed unmodified from /root/reference
    over NumPy stand-ins for the TensorFlow primitives (single-block scheme model and a two-block stack with a x2 up-sampling conv).
    The oracle must agree on the table index exactly and on the WaveNet output, the sub-band signals and the excitation to float32
    rounding (different summation order inside the convolutions)."""
    hp, plan, w = excitation_case(tag)
    orc = OracleMBExWN(hp, w, torch.float32)
    taps = {}
    exc = orc.generate_excitation(torch.as_tensor(EXC[f"{tag}_mel"]), torch.as_tensor(EXC[f"{tag}_f0"]),
                                  torch.as_tensor(EXC[f"{tag}_noise"]).reshape(2, -1, 1), taps).numpy()
    assert np.array_equal(taps["index"], EXC[f"{tag}_index"])
    for name, got in (("wn_out", taps["block_out_%d" % (len(plan.blocks) - 1)].numpy()), ("subbands", taps["subbands"].numpy()),
                      ("excitation", exc[:, :EXC[f"{tag}_excitation"].shape[1]])):
        ref = EXC[f"{tag}_{name}"]
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= 1e-5, (tag, name, err)


# ---- the whole inference branch: mel in, waveform out ------------------------------------------------------------------------
FWD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_forward.npz"))
_PQ = {"pulse_channels_use_pqmf": True, "pulse_channels_multi_band_config": {"subbands": 5, "taps": 40, "cutoff_ratio": 0.11, "beta": 8.0}}
FWD_CASES = {"speech": {}, "speech_lifter": {"ps_env_order_scale": 2.0},
             "band_gain_centered": {"ps_use_stft": False, "spect_filters_preserve_energy": True},
             "causal": {"force_causal": True},
             "pulse_pqmf_subharm": dict(_PQ, wavetable_config={"nominalF0": 60, "maxF0": 550, "add_subharm_chans": 1})}


def forward_case(tag):
    """(hparams, plan, weights) of a case of tests/golden/make_reference_forward_goldens.py."""
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(FWD_CASES[tag])
    plan = build_plan(hp)
    return hp, plan, W.init_synthetic(plan, seed=int(FWD[f"{tag}_seed"]))


@pytest.mark.parametrize("tag", sorted(FWD_CASES))
def test_oracle_forward_matches_the_reference_source(tag):
    """tests/golden/reference_forward.npz = MBExWN.call (inference branch) with generate_subnet_from_specs, generate_f0,
    generate_excitation, generate_specenv, _get_cepstral_windows and every layer `call` under them executed unmodified from
    /root/reference over NumPy stand-ins for the TensorFlow primitives (incl. tf.signal.stft / inverse_stft from their
    documentation).  F0, table index, lifter index, excitation, |VTF| and the waveform of the oracle must agree."""
    hp, plan, w = forward_case(tag)
    orc = OracleMBExWN(hp, w, torch.float32)
    r = orc.forward(FWD[f"{tag}_mel"], FWD[f"{tag}_noise"])
    f0 = FWD[f"{tag}_F0"]
    assert np.abs(r["F0"] - f0).max() <= 2e-5 * np.abs(f0).max()
    # the index is a discontinuous function of F0: compare it on the reference's own F0
    r = orc.forward(FWD[f"{tag}_mel"], FWD[f"{tag}_noise"], f0_override=f0)
    assert np.array_equal(r["index"], FWD[f"{tag}_index"])
    if FWD_CASES[tag].get("ps_env_order_scale"):
        assert np.array_equal(r["lifter_index"], FWD[f"{tag}_lifter_index"])
        assert len(np.unique(FWD[f"{tag}_lifter_index"])) > 3                   # the F0 contour selects several lifters
    for name, key, tol in (("excitation", "excitation", 1e-5), ("vtf_mag", "vtf", 1e-4), ("waveform", "waveform", 2e-5)):
        if f"{tag}_{name}" not in FWD.files:                                    # ps_use_stft = False has neither; short cases no |VTF|
            continue
        ref = FWD[f"{tag}_{name}"]
        got = np.abs(r[key]) if name == "vtf_mag" else r[key]
        got = got[:, :ref.shape[1]]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= tol, (tag, name, err)


# ---- NormMelComponents (row a3) ---------------------------------------------------------------------------------------------
NORM = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_norm.npz"))
NORM_CASES = {"default": {"normalize_rms_num_smooth_iters": 2},
              "one_iter_wide": {"normalize_rms_num_smooth_iters": 1, "normalize_smooth_win_scale": 2,
                                "normalize_smooth_with_squared_win": False, "max_norm_fact": 50.0, "normalize_compressor_exp": 0.8,
                                "use_max_limit": True},
              "pinv": {"normalize_rms_num_smooth_iters": 3, "normalize_use_pinv": True}}


@pytest.mark.parametrize("tag", sorted(NORM_CASES))
def test_oracle_norm_mel_matches_the_reference_source(tag):
    """tests/golden/reference_norm.npz = the reference's NormMelComponents (real constructor + normalize_inputs_by_rms,
    wavegen_1d.py:578-769) executed unmodified over NumPy stand-ins (tests/golden/make_reference_norm_goldens.py): the normalised
    log-mel and the sample-rate gain of oracle/norm_mel.py must agree, also when synth_length exceeds the smoothed gain (:762-766)."""
    from oracle.norm_mel import OracleNormMel
    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    orc = OracleNormMel(pc, NORM_CASES[tag])
    hop = pc["hop_size"]
    T = NORM["mell"].shape[1]
    for length, suffix in ((T * hop, ""), (T * hop + 170, "_long")):
        mell, rms, _ = orc.normalize_inputs_by_rms(NORM["mell"], length)
        ref_m, ref_r = NORM[f"{tag}{suffix}_mell"], NORM[f"{tag}{suffix}_rms"]
        assert mell.shape == ref_m.shape and rms.shape == ref_r.shape
        assert np.abs(mell - ref_m).max() <= 2e-5, (tag, suffix)                       # log domain: absolute
        assert np.abs(rms - ref_r).max() <= 2e-6 * np.abs(ref_r).max(), (tag, suffix)


# ---- the reference's model object, built by its own constructors -------------------------------------------------------------
MODEL_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_model.npz"))
MODEL_CASES = {"speech": {}, "blocks_2x1": EXC_CASES["blocks_2x1"], "lifter_causal": {"ps_env_order_scale": 1.5, "force_causal": True}}


@pytest.mark.parametrize("tag", sorted(MODEL_CASES))
def test_oracle_matches_the_reference_model_object(tag):
    """tests/golden/reference_model.npz: `MBExWN(preprocess_config=..., **mbexwn_config)` -- the reference's class with its own
    constructor (and those of WaveNetAEBlock / WaveNetAE / PulseWaveTable / TFPQMF), created from this package's config.yaml and
    run through PaNWaveNet.infer, all compiled unmodified from /root/reference over NumPy stand-ins
    (tests/golden/make_reference_model_goldens.py).  Pins what the other reference-source goldens leave to plan.py / dsp_init.py:
    rate algebra, dilation schedule, conditioning factors, lifter bank, F0 smoothing kernel."""
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(MODEL_CASES[tag])
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(MODEL_GOLD[f"{tag}_seed"]))
    orc = OracleMBExWN(hp, w, torch.float32)
    mel, nz, f0 = MODEL_GOLD[f"{tag}_mel"], MODEL_GOLD[f"{tag}_noise"], MODEL_GOLD[f"{tag}_F0"]
    r = orc.forward(mel, nz)
    assert np.abs(r["F0"] - f0).max() <= 2e-5 * np.abs(f0).max()
    r = orc.forward(mel, nz, f0_override=f0)
    assert np.array_equal(r["index"], MODEL_GOLD[f"{tag}_index"])
    wav = MODEL_GOLD[f"{tag}_waveform"]
    assert r["waveform"].shape == wav.shape
    assert np.abs(r["waveform"] - wav).max() <= 2e-5 * np.abs(wav).max(), tag
    if plan.env_order_scale:                                                    # the lifter bank of the reference's constructor
        assert np.array_equal(plan.lifters, MODEL_GOLD[f"{tag}_lifters"])
        assert np.array_equal(plan.lifter_log10f0, MODEL_GOLD[f"{tag}_lifter_grid"])


# ---- MELInverter.scale_mel (host side of the boundary) ----------------------------------------------------------------------
def test_scale_mel_matches_the_real_reference():
    """tests/golden/reference_scale_mel.npz = the reference's own MELInverter.scale_mel (mel_inverter.py:48-148, plain NumPy)
    called from /root/reference on the same inputs (tests/golden/make_reference_scale_mel_goldens.py): bit-identical."""
    import importlib.util
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    spec = importlib.util.spec_from_file_location(
        "make_scale_mel", os.path.join(os.path.dirname(__file__), "golden", "make_reference_scale_mel_goldens.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_scale_mel.npz"))
    for tag, (cfg, model_extra) in gen.cases().items():
        inv = MELInverter(None)
        m = dict(gen.MODEL, **model_extra)
        inv.hop_size, inv._srate, inv.fft_size, inv.fmin, inv.fmax, inv.mel_channels = (
            m["hop_size"], m["srate"], m["fft_size"], m["fmin"], m["fmax"], m["mel_channels"])
        inv.lin_amp_scale, inv.lin_amp_off, inv.mel_amp_scale, inv.use_max_limit = (
            m["lin_amp_scale"], m["lin_amp_off"], m["mel_amp_scale"], m["use_max_limit"])
        got = inv.scale_mel({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cfg.items()})
        assert got.dtype == np.float32 and got.shape == gold[tag].shape, tag
        assert np.array_equal(got, gold[tag]), tag


@pytest.mark.parametrize("tag", sorted(MODEL_CASES))
def test_checkpoint_paths_follow_the_reference_object_tree(tag, tmp_path):
    """A Keras checkpoint names a variable by the attribute path that leads to it.  reference_model.npz holds, for the reference's
    MBExWN object built by its own constructor, the position / kind / layer name of every member of pp_subnet_layers and
    ps_subnet_layers (pad, conv, PReLU and interpolation layers interleave) and the up_down_sample attribute of the WaveNet blocks.
    The checkpoint tf_checkpoint.export_weights writes must put each conv / PReLU variable at exactly that path, so that the
    unmodified reference can load_weights() it -- and import_weights must read a reference checkpoint the same way."""
    from mbexwn_vocoder_b200 import tf_checkpoint as T
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(MODEL_CASES[tag])
    plan = build_plan(hp, finalize=False)
    w = W.init_synthetic(plan, seed=int(MODEL_GOLD[f"{tag}_seed"]))
    prefix = str(tmp_path / "weights.tf")
    T.export_weights(prefix, hp, w)
    rd = T.BundleReader(prefix)
    keys = set(rd.keys())
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    n_conv = n_act = 0
    for row in MODEL_GOLD[f"{tag}_layer_table"]:
        path, kind, name = str(row).split("|")
        if kind == "conv":
            for part, leaf in (("v", "v"), ("g", "g"), ("bias", "conv1d_layer/bias")):
                key = f"block/{path}/{leaf}{suffix}"
                assert key in keys, key
                assert np.array_equal(np.asarray(rd.get(key)).reshape(w[f"{name}/{part}"].shape), w[f"{name}/{part}"]), key
            n_conv += 1
        elif kind == "prelu":
            key = f"block/{path}/alpha{suffix}"
            assert key in keys, key
            assert np.array_equal(np.asarray(rd.get(key)).reshape(-1), w[f"{name}/alpha"]), key
            n_act += 1
        elif kind == "block":
            has_up = f"block/{path}/up_down_sample/v{suffix}" in keys
            assert has_up == (name == "up_down_sample"), path
            assert f"block/{path}/wavenet/cond_layer/v{suffix}" in keys and f"block/{path}/wavenet/res_skip_layers/0/g{suffix}" in keys
    assert n_conv == sum(op.kind == "conv" for ops in (plan.pp_ops, plan.ps_ops) for op in ops) and n_act > 0
    assert f"block/wn_post_net/0/v{suffix}" in keys
