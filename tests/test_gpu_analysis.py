"""GPU parity of the analysis kernel (audio -> log-mel) through the C-ABI against the CPU oracle, which itself is pinned
on the real reference STFT (tests/test_analysis.py).  Tolerance: |STFT| and mel <= 1e-4 of the utterance peak (fp32)."""
import numpy as np
import pytest
import torch

from mbexwn_vocoder_b200 import get_config_file
from mbexwn_vocoder_b200.config import read_config

pytestmark = pytest.mark.gpu

LENGTHS = [4000, 700, 1, 299, 300, 24000, 1199, 601]          # ragged, shorter than the window, single sample, odd / even frame counts


@pytest.fixture(scope="module")
def setup():
    from mbexwn_vocoder_b200.analysis import MelAnalyzer
    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    return pc, MelAnalyzer(pc, device=0)


def _sounds():
    from oracle.analysis import synthetic_audio
    return [synthetic_audio(n, i) for i, n in enumerate(LENGTHS)]


def test_magnitude_and_mel_against_oracle(setup):
    from oracle import analysis as OA
    from mbexwn_vocoder_b200 import dsp_init as D
    pc, an = setup
    sounds = _sounds()
    mels, mags = an(sounds, do_post=False, return_mag=True)
    basis = D.mel_filter_bank(pc["sample_rate"], pc["fft_size"], pc["mel_channels"], pc["fmin"], pc["fmax"])
    for x, mel, mag in zip(sounds, mels, mags):
        S = OA.stft_magnitude(x[None], pc["win_size"], pc["hop_size"], pc["fft_size"])[0]
        assert mag.shape == S.shape == (x.size // 300 + 1, 1025)
        assert np.abs(mag - S).max() <= 1e-4 * S.max(), x.size
        ref_lin = S.astype(np.float64) @ basis.T.astype(np.float64)
        got_lin = np.exp(mel.astype(np.float64))
        floor = float(np.finfo(np.float32).eps)
        assert np.abs(got_lin - np.maximum(ref_lin, floor)).max() <= 1e-4 * ref_lin.max(), x.size
        ref = OA.compute_mel_spectrogram(x, pc, do_post=False)[0]
        assert mel.shape == ref.shape
        big = ref_lin > 1e-3 * ref_lin.max()
        assert np.abs(mel - ref)[big].max() <= 2e-3
        # the reference CLI's own figure of merit (bin/resynth_mel.py:92): mean |log-mel difference| in dB
        # (bins 100 dB below the peak carry fp32 FFT rounding noise, which the log amplifies)
        assert 8.685889638 * np.mean(np.abs(mel - ref)) < 0.1


def test_post_scaling_modes(setup):
    from oracle import analysis as OA
    from mbexwn_vocoder_b200.analysis import MelAnalyzer
    pc, an = setup
    sounds = _sounds()[:3]
    for cfg in (pc, dict(pc, use_max_limit=True, lin_amp_scale=2.0, mel_amp_scale=0.5, lin_amp_off=1e-4)):
        a = an if cfg is pc else MelAnalyzer(cfg, device=0)
        got = a(sounds, do_post=True)
        for x, m in zip(sounds, got):
            ref = OA.compute_mel_spectrogram(x, cfg, do_post=True)[0]
            # compare where the tolerance is defined: on the linear mel (1e-4 of the peak); in the log domain only
            # where the band energy is within 60 dB of the peak
            k = float(cfg.get("mel_amp_scale", 1))
            lin_got, lin_ref = np.exp(m.astype(np.float64) / k), np.exp(ref.astype(np.float64) / k)
            assert np.abs(lin_got - lin_ref).max() <= 1e-4 * lin_ref.max(), x.size
            big = lin_ref > 1e-3 * lin_ref.max()
            assert np.abs(m - ref)[big].max() <= 2e-3, x.size


def test_batch_composition_does_not_change_the_result(setup):
    pc, an = setup
    sounds = _sounds()
    together = an(sounds)
    for i in (0, 2, 5):
        alone = an([sounds[i]])[0]
        assert np.array_equal(alone, together[i])


def test_host_buffer_entry_point(setup):
    pc, an = setup
    sounds = _sounds()
    st = an.prepare([s.size for s in sounds])
    audio = torch.from_numpy(np.concatenate(sounds)).pin_memory()
    mel = torch.empty(st["mel"].shape, dtype=torch.float32).pin_memory()
    an.run_host(st, audio, mel)
    ref = np.concatenate(an(sounds))
    assert np.array_equal(mel.numpy(), ref)


def test_errors_are_loud(setup):
    import ctypes as C
    from mbexwn_vocoder_b200 import _cabi
    pc, an = setup
    st = an.prepare([600])
    cfg = an.config(False)
    cfg.fft_size = 1024
    rc = an.lib.mbexwn_mel_analysis(C.byref(cfg), C.byref(st["batch"]), None)
    assert rc == _cabi.ERR_UNSUPPORTED and b"2048" in an.lib.mbexwn_global_error()
    cfg = an.config(False)
    cfg.mode = 7
    assert an.lib.mbexwn_mel_analysis(C.byref(cfg), C.byref(st["batch"]), None) == _cabi.ERR_INVALID
    with pytest.raises(RuntimeError):
        an([np.zeros(0, np.float32)])


def test_resynthesis_round_trip_through_the_mel_inverter():
    """mel -> waveform -> mel with the reference-facing calls (bin/resynth_mel.py:80-96): shapes line up and the
    re-analysed mel is finite; with random weights the mel error itself is not meaningful."""
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    from oracle.forward import synthetic_mel
    inv = MELInverter("SPEECH", device=0, precision="f16f8")
    mell = np.stack([synthetic_mel(40, 3)])
    audio = inv.synth_from_mel(mell)
    dd = inv.generate_mel_from_snd(audio, srate=inv.srate)
    assert dd["mell"].shape == (80, 41) and dd["nfft"] == 2048 and dd["hoplen"] == 300 and dd["sr"] == 24000
    scaled = inv.scale_mel(dd)
    assert scaled.shape == (1, 41, 80) and np.all(np.isfinite(scaled))
    # a 2x oversampled copy of the waveform is brought back to the model rate on the host (sig_proc/resample.py)
    from mbexwn_vocoder_b200.analysis import resample
    up = resample(audio, 24000, 48000)
    dd2 = inv.generate_mel_from_snd(up, srate=48000)
    assert dd2["mell"].shape == (80, 41)
    # bands below 0.8 Nyquist (the two anti-aliasing filters take out the top of the spectrum)
    err_db = 8.685889638 * np.mean(np.abs(dd2["mell"][:70, 2:-2] - dd["mell"][:70, 2:-2]))
    assert err_db < 0.5
