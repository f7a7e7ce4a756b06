"""Init-time DSP (wavetables, LF pulse model, PQMF prototype) against vectors produced by the REAL reference code
(tests/golden/make_reference_goldens.py, generated in the build container where /root/reference is mounted)."""
import os

import numpy as np
import pytest

from mbexwn_vocoder_b200 import dsp_init

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_init_dsp.npz"))


@pytest.mark.parametrize("i", range(5))
def test_lf_synthesis_params_and_spectrum_match_reference(i):
    oq, am, ta, alpha, epar, ta_out = G[f"lfpar_{i}"]
    a, e, t = dsp_init.lf_synthesis_params(oq, am, ta)
    assert (a, e, t) == (alpha, epar, ta_out)
    f = G["lf_freqs"]
    with np.errstate(invalid="ignore", divide="ignore"):
        flow = dsp_init.lf_pulse_spectrum(f, oq, am, ta, derivative=False)
        deriv = dsp_init.lf_pulse_spectrum(f, oq, am, ta, derivative=True)
    assert np.array_equal(flow, G[f"lfspec_flow_{i}"], equal_nan=True)
    assert np.array_equal(deriv, G[f"lfspec_deriv_{i}"], equal_nan=True)


@pytest.mark.parametrize("tag", ["sp", "vo", "alt"])
def test_wavetable_bank_bit_exact_vs_reference(tag):
    sr, nominal, max_f0, realised = G[f"wt_{tag}_cfg"]
    wt = dsp_init.build_wavetables(sample_rate=sr, nominalF0=nominal, maxF0=max_f0)
    ref = G[f"wt_{tag}_tables"]
    assert wt.tables.dtype == np.float32 and wt.tables.shape == ref.shape
    assert np.array_equal(wt.tables, ref)
    assert wt.nominal_f0 == realised
    assert np.allclose(wt.f0_grid, G[f"wt_{tag}_grid"], rtol=0, atol=0)
    assert wt.n_period == ref.shape[0] - 1 and (wt.n_period & (wt.n_period - 1)) == 0
    assert np.array_equal(wt.tables[-1], wt.tables[0])          # wrap row (tf_wavetable.py:280)
    assert wt.tables.min() == -1.0                               # global normalisation by -min (tf_wavetable.py:286)


def test_wavetable_requires_max_f0_like_reference():
    with pytest.raises(TypeError):
        dsp_init.build_wavetables(sample_rate=8000.0, nominalF0=60.0)


@pytest.mark.parametrize("taps,cutoff,beta,key", [(240, 0.0377, 9.0, "pqmf_proto_240"), (62, 0.15, 9.0, "pqmf_proto_62")])
def test_pqmf_prototype_matches_reference(taps, cutoff, beta, key):
    assert np.array_equal(dsp_init.pqmf_prototype(taps, cutoff, beta), G[key])


def test_pqmf_polyphase_equals_dense_synthesis():
    S, taps = 15, 240
    _, syn = dsp_init.pqmf_filters(S, taps, 0.0377, 9.0)
    G3, Q, back = dsp_init.pqmf_polyphase(syn, S, taps)
    assert (Q, back) == (17, 8)
    rng = np.random.default_rng(0)
    T = 40
    x = rng.standard_normal((T, S))
    up = np.zeros((T * S + taps, S))
    up[taps // 2:taps // 2 + T * S:S] = x * S                   # zero-stuff with gain S, pad taps/2 both sides
    dense = np.array([np.sum(up[n:n + taps + 1] * syn.T.astype(np.float64)) for n in range(T * S)])
    xp = np.zeros((T + Q, S))
    xp[back:back + T] = x
    poly = np.einsum("mqk,qkp->mp", np.stack([xp[m:m + Q] for m in range(T)]), G3.astype(np.float64)).reshape(-1)
    assert np.abs(dense - poly).max() < 1e-6 * np.abs(dense).max()      # G is stored in float32


def test_pqmf_analysis_synthesis_reconstruction():
    """Near-perfect reconstruction of the cosine-modulated bank (property of tf_preprocess.py:120-145)."""
    S, taps = 15, 240
    ana, syn = dsp_init.pqmf_filters(S, taps, 0.0377, 9.0)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(S * 200)
    xp = np.pad(x, (taps // 2, taps // 2))
    sub = np.stack([np.correlate(xp, ana[k].astype(np.float64), mode="valid")[::S] for k in range(S)], axis=1)
    up = np.zeros((x.size + taps, S))
    up[taps // 2:taps // 2 + x.size:S] = sub * S
    y = sum(np.correlate(up[:, k], syn[k].astype(np.float64), mode="valid") for k in range(S))
    core = slice(taps, -taps)
    snr = 10 * np.log10(np.sum(x[core] ** 2) / np.sum((x[core] - y[core]) ** 2))
    assert snr > 40.0


def test_inverse_stft_window_is_hann_over_1p5():
    w = dsp_init.hann_periodic(1200)
    wi = dsp_init.inverse_stft_window(1200, 300)
    assert np.allclose(wi, w / 1.5, atol=1e-6)


def test_lifter_centre_tap_is_one():
    grid, lift = dsp_init.cepstral_lifters(1.0, 24000, 50.0, 550.0, 240)
    assert lift.shape == (30, 240) and np.all(lift[:, 0] == 1.0)    # custom_pulsed_generator.py:807
    assert np.all(np.diff(grid) > 0)
