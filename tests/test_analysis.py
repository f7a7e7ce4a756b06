"""Analysis side (audio -> log-mel) on the CPU: the oracle against vectors produced by the REAL reference
STFT / window code, the restated mel basis against its defining properties, host layout logic."""
import os

import numpy as np
import pytest

from mbexwn_vocoder_b200 import analysis as PA, dsp_init as D, get_config_file
from mbexwn_vocoder_b200.config import read_config
from oracle import analysis as OA

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_analysis.npz")


@pytest.fixture(scope="module")
def pc():
    return read_config(get_config_file("SPEECH"))["preprocess_config"]


def test_windows_equal_the_reference_generator():
    g = np.load(GOLD)
    assert np.array_equal(D.cosine_window("hann", 1200), g["hann1200"])
    assert np.array_equal(D.cosine_window("hann", 7), g["hann7"])
    assert np.array_equal(D.cosine_window("hamming", 8), g["hamming8"])
    with pytest.raises(RuntimeError, match="unsupported window"):
        D.cosine_window("gauss", 8)


def test_stft_magnitude_equals_the_reference_calc_stft():
    g = np.load(GOLD)
    S = OA.stft_magnitude(g["audio"], 1200, 300, 2048)
    assert S.dtype == np.float32 and S.shape == (2, 4000 // 300 + 1, 1025)
    assert np.array_equal(S, g["stft_mag"])
    # signal shorter than the window: np.pad(reflect) bounces more than once
    assert np.array_equal(OA.stft_magnitude(g["audio_short"], 1200, 300, 2048), g["stft_mag_short"])


def test_mel_basis_properties():
    sr, n_fft, n_mels = 24000, 2048, 80
    B = D.mel_filter_bank(sr, n_fft, n_mels, 0, 12000)
    assert B.shape == (n_mels, n_fft // 2 + 1) and B.dtype == np.float32 and B.min() >= 0
    freqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    # band edges equally spaced on the Slaney mel scale: linear (200/3 Hz per mel) below 1 kHz, log above
    edges = D._mel_to_hz_slaney(np.linspace(D._hz_to_mel_slaney(0), D._hz_to_mel_slaney(12000), n_mels + 2))
    assert np.isclose(D._hz_to_mel_slaney(1000.0), 15.0) and np.isclose(D._mel_to_hz_slaney(15.0), 1000.0)
    assert np.isclose(D._hz_to_mel_slaney(6400.0), 15.0 + 27.0)            # log step = log(6.4) / 27
    assert np.allclose(D._mel_to_hz_slaney(D._hz_to_mel_slaney(freqs)), freqs, atol=1e-6)
    for b in range(n_mels):
        nz = np.flatnonzero(B[b])
        assert freqs[nz[0]] > edges[b] - 1e-9 and freqs[nz[-1]] < edges[b + 2] + 1e-9      # support = (lo, hi)
        peak = freqs[np.argmax(B[b])]
        assert abs(peak - edges[b + 1]) <= sr / n_fft                                          # apex at the centre edge
        # Slaney normalisation: every triangle has unit area in Hz (to the bin-sampling error)
        assert abs(B[b].sum() * (sr / n_fft) - 1.0) < 0.02
    # neighbouring un-normalised triangles sum to one between the first and last centre
    U = D.mel_filter_bank(sr, n_fft, n_mels, 0, 12000, norm=False)
    inner = (freqs >= edges[1]) & (freqs <= edges[-2])
    assert np.allclose(U.sum(axis=0)[inner], 1.0, atol=1e-5)


def test_mel_filter_csr_reproduces_the_dense_product():
    B = D.mel_filter_bank(24000, 2048, 80, 0, 12000)
    lo, cnt, off, w = D.mel_filter_csr(B)
    assert w.size == int(cnt.sum()) and cnt.max() < 128
    S = np.random.default_rng(0).random(1025).astype(np.float32)
    dense = B @ S
    packed = np.array([np.dot(S[lo[b]:lo[b] + cnt[b]], w[off[b]:off[b] + cnt[b]]) for b in range(80)])
    assert np.allclose(packed, dense, rtol=1e-6, atol=1e-9)


def test_compute_mel_spectrogram_modes(pc):
    x = np.stack([OA.synthetic_audio(3000, 0), OA.synthetic_audio(3000, 1)])
    raw = OA.compute_mel_spectrogram(x, pc, do_post=False)
    assert raw.shape == (2, 11, 80) and raw.dtype == np.float32
    post = OA.compute_mel_spectrogram(x, pc, do_post=True)
    # do_post = log(mel + 1e-5) vs log(max(mel, eps)): equal where the mel energy is far above the offset
    big = raw > np.log(1e-2)
    assert big.any() and np.allclose(post[big], raw[big], atol=2e-3)
    pc2 = dict(pc, use_max_limit=True, lin_amp_scale=2.0, mel_amp_scale=0.5)
    lim = OA.compute_mel_spectrogram(x, pc2, do_post=True)
    assert np.allclose(lim, 0.5 * np.log(np.fmax(2.0 * np.exp(raw.astype(np.float64)), 1e-5)), atol=1e-5)


def test_analysis_layout():
    sb, ns, fb, pf = PA.analysis_layout([700, 4000, 300, 299], 300)
    assert sb.tolist() == [0, 700, 4700, 5000] and ns.tolist() == [700, 4000, 300, 299]
    assert fb.tolist() == [0, 3, 17, 19, 20]                                 # len // hop + 1 frames (stft.py:56)
    assert pf.tolist() == [0, 2, 9, 10, 11]                                  # one CTA per pair of frames
    assert PA.frame_count(4000, 300) == 14
    with pytest.raises(RuntimeError):
        PA.analysis_layout([10, 0], 300)


def test_resample_filter_follows_the_reference_design():
    fir, up, down = D.resample_filter(48000, 24000)
    assert (up, down) == (1, 2)
    # stop_att 70 dB => beta = 0.1102 (70 - 8.7); radius = ceil(62 / 2.285 / (2 pi 0.5 0.1) / 2) = 44 => 89 taps
    assert fir.size == 89 and np.isclose(fir.sum(), 1.0, atol=1e-6)
    assert np.allclose(fir, fir[::-1])
    fir2, up2, down2 = D.resample_filter(16000, 24000)
    assert (up2, down2) == (3, 2) and fir2.size % 3 == 0
    t = np.arange(4800) / 48000.0
    y = PA.resample(np.sin(2 * np.pi * 440.0 * t), 48000, 24000)
    assert y.size == 2400
    ref = np.sin(2 * np.pi * 440.0 * np.arange(2400) / 24000.0)
    assert np.abs(y[200:-200] - ref[200:-200]).max() < 2e-3


def test_no_cpu_fallback_for_analysis(pc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PA.MelAnalyzer(pc)


def test_oracle_analysis_matches_the_reference_functions():
    """`mell_*` of reference_analysis.npz = the reference's compute_mel_spectrogram_internal + scale_mel_spectrogram
    (preprocess.py:81-113, :417-572), AST-extracted and run unmodified with this package's mel basis standing in for librosa's:
    the oracle's framing, projection and log / scale post-processing must be bit-identical."""
    from mbexwn_vocoder_b200 import get_config_file
    from mbexwn_vocoder_b200.config import read_config
    from oracle.analysis import compute_mel_spectrogram
    g = np.load(GOLD)
    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    assert np.array_equal(compute_mel_spectrogram(g["audio"], pc), g["mell_post"])
    assert np.array_equal(compute_mel_spectrogram(g["audio"], pc, do_post=False), g["mell_nopost"])
    pc2 = dict(pc, lin_amp_scale=0.5, lin_amp_off=1e-3, mel_amp_scale=0.25, use_max_limit=True)
    assert np.array_equal(compute_mel_spectrogram(g["audio"], pc2), g["mell_post_scaled"])
