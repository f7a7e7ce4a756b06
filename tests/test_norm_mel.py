"""NormMelComponents (wavegen_1d.py:578-769): oracle closed forms + plan parsing on the CPU, GPU parity."""
import os

import numpy as np
import pytest
import yaml

from mbexwn_vocoder_b200 import get_config_file
from mbexwn_vocoder_b200.config import read_config
from mbexwn_vocoder_b200.plan import build_plan
from oracle.forward import synthetic_mel
from oracle.norm_mel import OracleNormMel, overlap_and_add

NORM_KEYS = {"normalize_rms_from_mell": True, "normalize_rms_num_smooth_iters": 2}


def _hp(extra=None):
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update(NORM_KEYS)
    hp["mbexwn_config"].update(extra or {})
    return hp


def test_overlap_and_add_closed_form():
    fr = np.arange(12, dtype=np.float32).reshape(1, 3, 4)
    out = overlap_and_add(fr, 2)
    assert out.tolist() == [[0, 1, 2 + 4, 3 + 5, 6 + 8, 7 + 9, 10, 11]]


def test_constant_rms_is_a_fixed_point_and_mel_is_divided_by_it():
    hp = _hp()
    o = OracleNormMel(hp["preprocess_config"], hp["mbexwn_config"])
    mell = np.tile(synthetic_mel(1, 0), (1, 30, 1))
    out, up, taps = o.normalize_inputs_by_rms(mell, 30 * 300)
    r = taps["rms_raw"][0, 0]
    assert np.allclose(taps["rms"], r, rtol=1e-5) and np.allclose(up, r, rtol=1e-5)
    assert np.allclose(out, np.log(np.exp(mell) / r + 1e-5), atol=1e-5)
    # raw estimate: sqrt(sum (mel * (f_hi - f_lo) / 2)^2 / (fft * win / 2))   (wavegen_1d.py:598, :611, :690)
    from mbexwn_vocoder_b200 import dsp_init as D
    f = D.mel_frequencies(82, 0, 12000)
    expect = np.sqrt(np.sum((np.exp(mell[0, 0]) * (f[2:] - f[:-2]) / 2) ** 2) / (2048 * 1200 * 0.5))
    assert np.isclose(r, expect, rtol=1e-5)


def test_smoothing_is_linear_and_preserves_scale():
    hp = _hp({"normalize_rms_num_smooth_iters": 3})
    o = OracleNormMel(hp["preprocess_config"], hp["mbexwn_config"])
    mell = synthetic_mel(40, 2)[None]
    _, up1, t1 = o.normalize_inputs_by_rms(mell, 40 * 300)
    _, up2, t2 = o.normalize_inputs_by_rms(mell + np.log(3.0).astype(np.float32), 40 * 300)     # mel x 3 => rms x 3
    assert np.allclose(t2["rms"], 3 * t1["rms"], rtol=1e-4) and np.allclose(up2, 3 * up1, rtol=1e-4)
    assert t1["rms"].min() >= t1["rms_raw"].min() * 0.999 and t1["rms"].max() <= t1["rms_raw"].max() * 1.001
    assert np.abs(np.diff(t1["rms"][0])).max() < np.abs(np.diff(t1["rms_raw"][0])).max()         # smoother


def test_options_floor_compressor_pinv_and_limits():
    hp = _hp({"max_norm_fact": 10.0, "normalize_compressor_exp": 0.5, "use_max_limit": True, "lin_amp_scale": 2.0,
              "mel_amp_scale": 0.5, "lin_amp_off": 1e-4})
    o = OracleNormMel(hp["preprocess_config"], hp["mbexwn_config"])
    mell = synthetic_mel(12, 1)[None] - 6.0                      # quiet: the floor 1 / max_norm_fact is active
    out, up, taps = o.normalize_inputs_by_rms(mell.astype(np.float32), 12 * 300)
    assert np.allclose(taps["rms_raw"], np.sqrt(0.1), rtol=1e-6)
    assert np.allclose(out, 0.5 * np.log(np.maximum(np.exp(mell) / taps["rms"][:, :, None] * 2.0, 1e-4)), atol=1e-5)
    hp2 = _hp({"normalize_use_pinv": True})
    o2 = OracleNormMel(hp2["preprocess_config"], hp2["mbexwn_config"])
    _, _, t2 = o2.normalize_inputs_by_rms(synthetic_mel(12, 1)[None], 12 * 300)
    assert t2["rms_raw"].shape == (1, 12) and np.all(t2["rms_raw"] > 0)
    with pytest.raises(NotImplementedError):
        OracleNormMel(hp["preprocess_config"], {"normalize_rms_num_smooth_iters": 0})
    with pytest.raises(RuntimeError, match="4 \\* hop_size"):
        OracleNormMel(dict(hp["preprocess_config"], win_size=1024), hp["mbexwn_config"])


def test_plan_parses_the_normaliser_keys():
    plan = build_plan(_hp({"max_norm_fact": 20.0, "normalize_smooth_win_scale": 2}))
    nm = plan.norm
    assert nm is not None and nm.iters == 2 and nm.win == 1200 and nm.smooth_win == 2400 and nm.squared_win
    assert np.isclose(nm.floor, 0.05) and nm.norm_fact == 2048 * 1200 * 0.5 and nm.proj.shape == (80,) and nm.proj_cols == 0
    assert nm.smooth_window.shape == (2400,) and np.isclose(nm.gwin.sum(), 1.0, atol=1e-6)
    pinv = build_plan(_hp({"normalize_use_pinv": True})).norm
    assert pinv.proj.shape == (80, 1025) and pinv.proj_cols == 1025 and pinv.proj_scale < 0.1
    assert build_plan(read_config(get_config_file("SPEECH"))).norm is None
    with pytest.raises(NotImplementedError):
        build_plan(_hp({"normalize_rms_num_smooth_iters": 0}))


def _model_dir(tmp_path, extra):
    """A model directory = the SPEECH config plus normaliser keys."""
    cfg = yaml.safe_load(open(get_config_file("SPEECH")))
    cfg["mbexwn_config"].update(NORM_KEYS)
    cfg["mbexwn_config"].update(extra)
    d = tmp_path / ("m" + str(abs(hash(str(sorted(extra.items())))) % 10 ** 8))
    os.makedirs(d, exist_ok=True)
    yaml.safe_dump(cfg, open(d / "config.yaml", "w"))
    return str(d)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [{}, {"normalize_rms_num_smooth_iters": 1, "normalize_smooth_win_scale": 2,
                                        "normalize_smooth_with_squared_win": False, "max_norm_fact": 50.0,
                                        "normalize_compressor_exp": 0.8, "use_max_limit": True, "lin_amp_scale": 1.5},
                                   {"normalize_use_pinv": True}])
def test_gpu_normaliser_against_oracle(tmp_path, extra):
    """mel_norm / rms / gain taps and the final waveform with the normaliser on, against oracle(normalise) -> oracle forward
    -> x upsampled_rms, ragged batch, fp32 and the tensor-core path."""
    import torch
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    from oracle.forward import OracleMBExWN, synthetic_noise
    inv = MELInverter(_model_dir(tmp_path, extra), device=0, precision="fp32")
    plan = inv.plan
    hp = read_config(inv.config_file)
    onorm = OracleNormMel(hp["preprocess_config"], hp["mbexwn_config"])
    hp_plain = read_config(inv.config_file)
    for k in list(hp_plain["mbexwn_config"]):
        if k.startswith("normalize_"):
            hp_plain["mbexwn_config"].pop(k)
    oracle = OracleMBExWN(hp_plain, inv.weights, torch.float32)
    lengths = [24, 7, 13]
    mels = [synthetic_mel(t, i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, i) for i, t in enumerate(lengths)]
    refs = []
    for m, z, t in zip(mels, noise, lengths):
        nm, up, taps = onorm.normalize_inputs_by_rms(m[None], t * plan.hop)
        f0 = oracle.generate_f0(torch.as_tensor(nm)).numpy()
        r = oracle.forward(nm, z[None], f0_override=f0)
        refs.append((nm[0], up[0], taps, f0[0], r["waveform"][0] * up[0]))
    for precision in ("fp32", "f16f8"):
        inv.precision = precision
        out, taps = inv.synth_batch(mels, noise=noise, f0=[r[3] for r in refs], taps=["mel_norm", "norm_gain"])
        for u in range(len(lengths)):
            nm, up, otaps, _, wav = refs[u]
            assert np.abs(taps["mel_norm"][u].reshape(nm.shape) - nm).max() <= 1e-4 * np.abs(nm).max(), (precision, u)
            g = taps["norm_gain"][u].reshape(-1)
            assert np.abs(g - up).max() <= 1e-5 * up.max(), (precision, u)
            err = out[u].astype(np.float64) - wav
            snr = 10 * np.log10(np.sum(wav.astype(np.float64) ** 2) / max(np.sum(err ** 2), 1e-300))
            assert snr >= 60.0, (precision, u, snr)
    with pytest.raises(NotImplementedError):
        inv.synth_long_from_mel(mels[0])
