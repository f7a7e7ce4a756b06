"""Variants of the path the scheme configuration does not use: the spectral-shaping stage without the STFT
filter (custom_pulsed_generator.py:666-674: ps_use_stft = False with the per-band gain of :857-884, ps_off), sub-harmonic
channels, PQMF analysis of the pulse train, force_causal and multi-block / up-sampling WaveNet stacks (:459-488,
custom_AE_layers.py:457-575) -- on the CPU (plan / oracle) and against the oracle on the GPU."""
import os

import numpy as np
import pytest
import torch
import yaml

from mbexwn_vocoder_b200 import get_config_file, tf_checkpoint as T, weights as W
from mbexwn_vocoder_b200.config import read_config
from mbexwn_vocoder_b200.plan import PS_BAND_GAIN, PS_OFF, PS_STFT, build_plan
from oracle.forward import OracleMBExWN, lin_interp, synthetic_mel, synthetic_noise

_PULSE_PQMF = {"pulse_channels_use_pqmf": True,
               "pulse_channels_multi_band_config": {"subbands": 5, "taps": 40, "cutoff_ratio": 0.11, "beta": 8.0}}
VARIANTS = {"causal": {"force_causal": True},
            "pulse_pqmf": dict(_PULSE_PQMF),
            "pulse_pqmf_subharm": dict(_PULSE_PQMF, wavetable_config={"nominalF0": 60, "maxF0": 550, "add_subharm_chans": 1}),
            "subharm": {"wavetable_config": {"nominalF0": 60, "maxF0": 550, "add_subharm_chans": 2}},
            "band_gain": {"ps_use_stft": False}, "band_gain_centered": {"ps_use_stft": False, "spect_filters_preserve_energy": True},
            "ps_off": {"ps_off": True},
            # pp_waveNetBlocks: a block at 800 Hz with a x2 sub-pixel up-sampling conv, then a block at 1600 Hz ...
            "blocks_2x1": {"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                           "pp_mod_subnet_channel_factors": [0.5, 0.25]},
            # ... two blocks at 800 Hz, the up-sampling conv behind the last one feeds the post net
            "blocks_1x2": {"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [1, 2],
                           "pp_mod_subnet_channel_factors": [0.25, 0.5]}}


def _wavenet_with(**kw):
    sub = dict(yaml.safe_load(open(get_config_file("SPEECH")))["mbexwn_config"]["pp_mod_subnet"])
    sub.update(kw)
    return {"pp_mod_subnet": sub}


# the other gated units of WaveNetAE (custom_AE_layers.py:273-304): the fused layer kernel compiles tanh * sigmoid in and keeps
# a run-time switch for these
VARIANTS.update({"gate_gfu": _wavenet_with(activation="gfu"), "gate_gsu": _wavenet_with(activation="gsu"),
                 "gate_glu": _wavenet_with(activation="glu")})


def _hp(extra):
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(extra)
    return hp


def test_plan_modes_and_subnet_width():
    assert build_plan(_hp({}), finalize=False).ps_mode == PS_STFT
    p = build_plan(_hp(VARIANTS["band_gain"]), finalize=False)
    assert p.ps_mode == PS_BAND_GAIN and p.ps_ops[-1].conv.cout == p.subbands == 15
    assert build_plan(_hp(VARIANTS["band_gain_centered"]), finalize=False).ps_preserve_energy
    q = build_plan(_hp(VARIANTS["ps_off"]), finalize=False)
    assert q.ps_mode == PS_OFF and q.ps_ops == [] and not any(l.name.startswith("PS_") for l in q.conv_layers())
    with pytest.raises(NotImplementedError):
        build_plan(_hp({"spect_filters_preserve_energy": True}), finalize=False)


def test_subharmonic_channels_widen_the_wavenet_input():
    """wavetable_config.add_subharm_chans = n: every pulse sample brings sin(2 pi phase / ii), ii = 2 .. n + 1, along
    (tf_wavetable.py:554-559); rows are [p0 s0,2 s0,3 p1 s1,2 ... | noise] (custom_pulsed_generator.py:893)."""
    hp = _hp(VARIANTS["subharm"])
    plan = build_plan(hp)
    assert plan.subharm == 2 and plan.wavenet.c_in == 5 * 3 + 1
    assert [l for l in plan.conv_layers() if l.name.endswith("/start")][0].cin == 16
    orc = OracleMBExWN(hp, W.init_synthetic(plan, seed=12), torch.float32)
    mel = synthetic_mel(6, 0)[None]
    r = orc.forward(mel, synthetic_noise(6 * plan.steps_per_frame, 0)[None])
    x = r["wn_in"][0]
    assert x.shape == (6 * 20, 16)
    ph = r["phase"][0].reshape(-1, 5)
    assert np.allclose(x[:, 0::3][:, :5], r["pulse"][0].reshape(-1, 5))
    assert np.allclose(x[:, 1::3][:, :5], np.sin(2 * np.pi * ph / 2), atol=1e-6)
    assert np.allclose(x[:, 2::3][:, :5], np.sin(2 * np.pi * ph / 3), atol=1e-6)


def test_force_causal_moves_every_pad_to_the_left():
    """force_causal (custom_pulsed_generator.py:53, :76-81, :474-475): no output sample depends on a later mel frame, up to
    the look-ahead of the stages that are not convolutions (LinInterp's next row, PQMF, STFT)."""
    hp = _hp(VARIANTS["causal"])
    plan = build_plan(hp)
    for layer in plan.conv_layers():
        assert layer.pad_r == 0 and layer.pad_l == (layer.k - 1) * layer.dilation, layer.name
    assert plan.wavenet.causal
    orc = OracleMBExWN(hp, W.init_synthetic(plan, seed=12), torch.float32)
    T = 30
    mel = synthetic_mel(T, 0)[None]
    mel2 = mel.copy()
    mel2[:, 20:] = synthetic_mel(T, 5)[None][:, 20:]            # change the future of frame 20
    nz = synthetic_noise(T * plan.steps_per_frame, 0)[None]
    a, b = orc.forward(mel, nz), orc.forward(mel2, nz)
    # WaveNet rows strictly before frame 19 (one frame of LinInterp look-ahead in the conditioning) are unchanged
    assert np.array_equal(a["wn_out"][0, :18 * 20], b["wn_out"][0, :18 * 20])
    assert not np.array_equal(a["wn_out"][0, 21 * 20:], b["wn_out"][0, 21 * 20:])
    # the non-causal model looks ahead: the same experiment changes earlier rows
    hp0 = _hp({})
    orc0 = OracleMBExWN(hp0, W.init_synthetic(build_plan(hp0), seed=12), torch.float32)
    a0, b0 = orc0.forward(mel, nz), orc0.forward(mel2, nz)
    assert not np.array_equal(a0["wn_out"][0, :18 * 20], b0["wn_out"][0, :18 * 20])


def test_pulse_pqmf_analysis_input():
    """pulse_channels_use_pqmf: the 5 WaveNet pulse channels are the bands of a PQMF analysis of the pulse train (decimated by
    5) instead of 5 consecutive samples; an analysis -> synthesis round trip of the bank reconstructs the pulse."""
    hp = _hp(VARIANTS["pulse_pqmf"])
    plan = build_plan(hp)
    assert plan.pulse_pqmf_ana.shape == (5, 41) and plan.wavenet.c_in == 6
    with pytest.raises(RuntimeError, match="pulse_channels"):
        build_plan(_hp(dict(_PULSE_PQMF, pulse_channels_multi_band_config={"subbands": 4, "taps": 40, "cutoff_ratio": 0.11, "beta": 8.0})))
    orc = OracleMBExWN(hp, W.init_synthetic(plan, seed=12), torch.float32)
    mel = synthetic_mel(8, 0)[None]
    r = orc.forward(mel, synthetic_noise(8 * plan.steps_per_frame, 0)[None])
    x, pulse = r["wn_in"][0], r["pulse"][0]
    assert x.shape == (8 * 20, 6)
    ana = plan.pulse_pqmf_ana
    m, k = 37, 3                                              # one value by hand: sum_j pulse[5 m + j - 20] h_k[j]
    idx = 5 * m + np.arange(41) - 20
    ok = (idx >= 0) & (idx < pulse.size)
    assert np.isclose(x[m, k], np.sum(pulse[idx[ok]] * ana[k][ok]), atol=1e-5)


def test_multi_block_wavenet_plan_and_oracle():
    """pp_mod_subnet_upsampling_factors / _channel_factors build one WaveNetAEBlock each (custom_pulsed_generator.py:465-488):
    rates multiply by the up-sampling factors, every block has its own conditioning conv (sub-pixel factor = block rate /
    (frame rate x cond_lin_upsampling)), blocks behind the first read n_out_channels, and a block with factor > 1 ends in a
    k = 3 sub-pixel conv named <block>_WNBlock_UP_<factor> (custom_AE_layers.py:519-526)."""
    hp = _hp(VARIANTS["blocks_2x1"])
    plan = build_plan(hp)
    b0, b1 = plan.blocks
    assert plan.wavenet is b0 and (plan.steps_per_frame, plan.sub_per_frame, plan.pulse_per_frame) == (10, 20, 100)
    assert (b0.name, b0.c, b0.c_in, b0.steps_per_frame, b0.cond_conv_up, b0.up) == ("PP_waveNetBlock_ups2_0", 160, 11, 10, 1, 2)
    assert (b1.name, b1.c, b1.c_in, b1.steps_per_frame, b1.cond_conv_up, b1.up) == ("PP_waveNetBlock_ups1_1", 80, 30, 20, 2, 1)
    ups = [l for l in plan.conv_layers() if "_WNBlock_UP_" in l.name]
    assert [(l.name, l.k, l.cin, l.cout, l.subpixel) for l in ups] == [("PP_waveNetBlock_ups2_0_WNBlock_UP_2", 3, 30, 60, 2)]
    single = build_plan(_hp({}), finalize=False)
    assert len(single.blocks) == 1 and single.blocks[0].up == 1 and single.sub_per_frame == single.steps_per_frame == 20
    with pytest.raises(RuntimeError, match="generated sample rate"):                 # custom_pulsed_generator.py:344
        build_plan(_hp({"pp_mod_subnet_upsampling_factors": [2, 1], "pp_mod_subnet_channel_factors": [1, 1]}), finalize=False)
    with pytest.raises(RuntimeError, match="cannot achieve conditioning rate"):      # :469: 400 Hz is not 80 Hz x 10 x n
        build_plan(_hp({"pulse_channels": 20, "pp_mod_subnet_upsampling_factors": [4], "pp_mod_subnet_channel_factors": [1]}),
                   finalize=False)
    w = W.init_synthetic(plan, seed=12)
    orc = OracleMBExWN(hp, w, torch.float32)
    T_ = 7
    mel = synthetic_mel(T_, 0)[None]
    nz = synthetic_noise(T_ * plan.steps_per_frame, 0)[None]
    r = orc.forward(mel, nz)
    assert r["wn_in"].shape == (1, T_ * 10, 11) and r["block_out_0"].shape == (1, T_ * 20, 30)
    assert r["wn_out"].shape == (1, T_ * 20, 30) and r["subbands"].shape == (1, T_ * 20, 15) and r["waveform"].shape == (1, T_ * 300)
    # the up-sampling conv by hand: row 2 t + s of its output = channels [30 s, 30 s + 30) of the k = 3 SAME conv at row t
    x0 = orc.wavenet_block(torch.as_tensor(r["wn_in"]), torch.as_tensor(mel), 0)
    kern, bias = W.folded(w, b0.up_name)
    t = 4
    row = sum(x0[0, t - 1 + j].numpy() @ kern[j] for j in range(3)) + bias
    assert np.allclose(r["block_out_0"][0, 2 * t], row[:30], atol=1e-5) and np.allclose(r["block_out_0"][0, 2 * t + 1], row[30:], atol=1e-5)
    # the up-sampling conv behind the last block feeds the post net
    hp2 = _hp(VARIANTS["blocks_1x2"])
    plan2 = build_plan(hp2)
    assert [(b.steps_per_frame, b.cond_conv_up, b.up) for b in plan2.blocks] == [(10, 1, 1), (10, 1, 2)] and plan2.sub_per_frame == 20
    r2 = OracleMBExWN(hp2, W.init_synthetic(plan2, seed=12), torch.float32).forward(mel, nz)
    assert r2["wn_out"].shape == (1, T_ * 10, 30) and r2["block_out_1"].shape == (1, T_ * 20, 30) and r2["subbands"].shape == (1, T_ * 20, 15)


def test_multi_block_cabi_config_and_context():
    from mbexwn_vocoder_b200.engine import make_config
    from mbexwn_vocoder_b200.long_form import main_context_frames
    plan = build_plan(_hp(VARIANTS["blocks_2x1"]))
    c = make_config(plan)
    assert c.wn_n_blocks == 2 and c.steps_per_frame == 10 and c.wn_cin == 11 and c.wn_cout == 30
    assert [(b.c, b.cond_conv_up, b.up) for b in c.wn_blocks[:2]] == [(160, 1, 2), (80, 2, 1)]
    assert c.wn_blocks[0].up_name == b"PP_waveNetBlock_ups2_0_WNBlock_UP_2" and c.wn_blocks[1].name == b"PP_waveNetBlock_ups1_1_WNBlock_WN"
    assert make_config(build_plan(_hp({}))).wn_n_blocks == 0
    # receptive fields of the blocks add up: 30 rows at 10 per frame + 2 (up conv) and 30 rows at 20 per frame
    assert main_context_frames(plan) >= 4 + 2 + main_context_frames(build_plan(_hp({}))) - 2 - 1


def test_checkpoint_round_trip_of_the_variants(tmp_path):
    for name, extra in VARIANTS.items():
        hp = _hp(extra)
        plan = build_plan(hp, finalize=False)
        w = W.init_synthetic(plan, seed=2)
        prefix = str(tmp_path / name / "weights.tf")
        T.export_weights(prefix, hp, w)
        got = T.import_weights(prefix, plan)
        assert sorted(got) == sorted(w) and all(np.array_equal(got[k], w[k]) for k in w)


def test_oracle_band_gain_semantics():
    """The gain sequence is interpolated to the sample rate (x hop) but multiplies the sub-band rows: row r of an utterance
    reads value r of that sequence (custom_pulsed_generator.py:453, :917); ps_off = the bare PQMF output."""
    hp = _hp(VARIANTS["band_gain"])
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=12)
    orc = OracleMBExWN(hp, w, torch.float32)
    mel = synthetic_mel(12, 0)[None]
    nz = synthetic_noise(12 * plan.steps_per_frame, 0)[None]
    r = orc.forward(mel, nz)
    assert r["waveform"].shape == (1, 12 * 300) and "vtf" not in r
    g = orc.generate_multiband_gain(torch.as_tensor(mel))
    full = lin_interp(g, 300, num_pad_end=1, drop_last=False)
    assert full.shape == (1, 12 * 300 + 1, 15)
    assert np.allclose(r["mb_gain"], full[:, :12 * 20].numpy())
    assert np.allclose(r["mb_gain"][0, 0], g[0, 0].numpy()) and np.all(r["mb_gain"] > 0)
    # frame 0 -> frame 1 is reached only after 300 rows, i.e. never within a 12-frame (240-row) utterance
    assert np.allclose(r["mb_gain"][0, 150], 0.5 * (g[0, 0] + g[0, 1]).numpy(), rtol=1e-5)
    hp_off = _hp(VARIANTS["ps_off"])
    plan_off = build_plan(hp_off)
    orc_off = OracleMBExWN(hp_off, W.init_synthetic(plan_off, seed=12), torch.float32)
    r_off = orc_off.forward(mel, nz)
    assert np.array_equal(r_off["waveform"], r_off["excitation"][:, :12 * 300])


def _model_dir(tmp_path, name):
    cfg = yaml.safe_load(open(get_config_file("SPEECH")))
    cfg["mbexwn_config"].update(VARIANTS[name])
    d = tmp_path / name
    os.makedirs(d, exist_ok=True)
    yaml.safe_dump(cfg, open(d / "config.yaml", "w"))
    return str(d)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_gpu_variants_against_oracle(tmp_path, name):
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    inv = MELInverter(_model_dir(tmp_path, name), device=0, precision="fp32")
    plan = inv.plan
    oracle = OracleMBExWN(read_config(inv.config_file), inv.weights, torch.float32)
    # (tf.pad SYMMETRIC cannot mirror 2 rows out of a 1-frame utterance: the causal reference needs >= 2 frames)
    lengths = [33, 9, 2] if name == "causal" else [33, 9, 1]
    mels = [synthetic_mel(t, i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, i) for i, t in enumerate(lengths)]
    f0 = [oracle.generate_f0(torch.as_tensor(m[None])).numpy()[0] for m in mels]
    refs = [oracle.forward(m[None], z[None], f0_override=f[None]) for m, z, f in zip(mels, noise, f0)]
    for precision in ("fp32", "f16f8"):
        inv.precision = precision
        out, taps = inv.synth_batch(mels, noise=noise, f0=f0, taps=["subbands", "index"])
        for u in range(len(lengths)):
            ref = refs[u]
            assert np.array_equal(taps["index"][u], ref["index"][0].reshape(-1))
            if name in ("subharm", "pulse_pqmf", "pulse_pqmf_subharm"):
                x = ref["wn_in"][0]
                got = inv.synth_batch(mels, noise=noise, f0=f0, taps=["wn_in"])[1]["wn_in"][u].reshape(x.shape)
                assert np.abs(got - x).max() <= 1e-5 * np.abs(x).max()
            sub = ref["subbands"][0]                           # after the band gain, i.e. what enters the PQMF
            assert np.abs(taps["subbands"][u].reshape(sub.shape) - sub).max() <= 1e-4 * np.abs(sub).max(), (precision, u)
            wav = ref["waveform"][0].astype(np.float64)
            err = out[u].astype(np.float64) - wav
            snr = 10 * np.log10(np.sum(wav ** 2) / max(np.sum(err ** 2), 1e-300))
            assert snr >= 60.0, (precision, u, snr)
    if name in ("causal", "blocks_2x1"):
        # chunked long-form synthesis stays bit-identical with the one-sided receptive field / the blocks' summed context
        inv.precision = "f16f8"
        T = 131
        mel = synthetic_mel(T, 9)
        nz = synthetic_noise(T * plan.steps_per_frame, 9).reshape(-1)
        whole = inv.synth_from_mel(mel[None], noise=[nz])
        chunked = inv.synth_long_from_mel(mel, noise=nz, chunk_frames=40)
        assert np.array_equal(whole, chunked)
