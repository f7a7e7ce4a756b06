"""Chunked long-form synthesis (BASELINE.json configs[4]) must reproduce the un-chunked forward sample for sample."""
import numpy as np
import pytest
import torch

from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def inverter():
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    inv = MELInverter("SPEECH", device=0, precision="f16f8")
    yield inv
    inv.model.close()


@pytest.mark.parametrize("frames,chunk", [(173, 40), (96, 25), (61, 100)])
def test_chunked_equals_unchunked(inverter, frames, chunk):
    plan = inverter.plan
    mel = synthetic_mel(frames, 21)
    noise = synthetic_noise(frames * plan.steps_per_frame, 21).reshape(-1)
    whole = inverter.synth_from_mel(mel[None], noise=[noise])
    for max_batch in (32768, 300):                 # all windows in one call / several calls
        parts, info = inverter.synth_long_from_mel(mel, noise=noise, chunk_frames=chunk, max_batch_frames=max_batch,
                                                   return_info=True)
        assert parts.shape == whole.shape
        assert info["n_windows"] == -(-frames // chunk)
        # every kernel is row-local with a fixed summation order: the cores are bit-identical
        assert np.array_equal(parts, whole), np.abs(parts - whole).max()


def test_chunked_matches_oracle(inverter, speech_setup):
    """... and therefore the oracle, at the tolerance of the fp32-accurate path."""
    hp, plan, w = speech_setup
    oracle = OracleMBExWN(hp, w, torch.float32)
    frames = 120
    mel = synthetic_mel(frames, 3)
    noise = synthetic_noise(frames * plan.steps_per_frame, 3)
    ref = oracle.forward(mel[None], noise[None])["waveform"][0].astype(np.float64)
    out = inverter.synth_long_from_mel(mel, noise=noise.reshape(-1), chunk_frames=30).astype(np.float64)
    snr = 10 * np.log10(np.sum(ref ** 2) / np.sum((out - ref) ** 2))
    assert snr >= 60.0


def test_phase_carry_host_restatement_matches_kernels(inverter):
    """phase_run_before_chunks (host) against the device scan: the 'phase' tap of a window started with the carry equals
    the tap of the whole utterance."""
    from mbexwn_vocoder_b200.long_form import phase_run_before_chunks
    plan = inverter.plan
    frames, cut = 90, 40                      # window starts at frame 40 = chunk 4
    n = frames * plan.pulse_per_frame
    f0 = (110.0 * 2 ** (np.sin(np.arange(n) / 900.0))).astype(np.float32)
    mel = synthetic_mel(frames, 5)
    eng = inverter.model
    _, tp = eng.forward([mel], f0=[f0], precision="f16f8", taps=["phase", "index"])
    run = phase_run_before_chunks(f0, plan.pulse_rate)
    pb = eng.prepare([frames - cut], "f16f8", with_noise=False, with_f0=True, with_carry=True)
    pb.load([mel[cut:]], f0=[f0[cut * plan.pulse_per_frame:]], carry=[run[cut * plan.pulse_per_frame // 1000]])
    pb.run_host()
    assert np.array_equal(pb.tap("phase")[0], tp["phase"][0][cut * plan.pulse_per_frame:])
    assert np.array_equal(pb.tap("index")[0], tp["index"][0][cut * plan.pulse_per_frame:])


def test_device_phase_carry_equals_the_host_restatement(inverter):
    """mbexwn_phase_carry (device: chunk totals + sequential scan) against phase_run_before_chunks (NumPy restatement of
    tf_wavetable.py:470-486) on a ten-minute F0 track: bit for bit, all 4800 chunks."""
    import ctypes as C
    from mbexwn_vocoder_b200 import _cabi
    from mbexwn_vocoder_b200.long_form import phase_run_before_chunks
    plan, eng = inverter.plan, inverter.model
    n = 48000 * plan.pulse_per_frame
    t = np.arange(n, dtype=np.float64)
    f0 = (180.0 * 2 ** (0.8 * np.sin(t / 7000.0) + 0.1 * np.sin(t / 311.0))).astype(np.float32)
    ref = phase_run_before_chunks(f0, plan.pulse_rate)
    f0_dev = torch.from_numpy(f0).cuda()
    run = torch.empty(ref.size, dtype=torch.float32, device="cuda")
    rc = eng.lib.mbexwn_phase_carry(eng._handle, f0_dev.data_ptr(), n, run.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(eng.lib, eng._handle, rc, "mbexwn_phase_carry")
    assert np.array_equal(run.cpu().numpy(), ref)


def test_ten_minute_mel_chunked_equals_unchunked(inverter, speech_setup):
    """BASELINE.json configs[4] at its full size: one 48 000-frame (10-minute) mel.  Chunked synthesis (400-frame cores,
    device-resident F0 track and phase carry: 4800 unwrapped float32 chunk offsets) against ONE forward over the whole
    signal: bit for bit over all 14.4 M samples, for the device-resident form and for the host-loop form.  Against the oracle:
    the wrapped phase of the whole signal (4.8 M samples behind 4800 chunk offsets) bit for bit, and a window 8 s into the
    signal at 1e-4 of peak / 60 dB (the oracle runs the 12 s prefix)."""
    hp, plan, w = speech_setup
    T = 48000
    ppf, spf, hop = plan.pulse_per_frame, plan.steps_per_frame, plan.hop
    mel = np.concatenate([synthetic_mel(2400, 900 + k) for k in range(T // 2400)])
    noise = np.random.default_rng(5).standard_normal(T * spf, dtype=np.float32)
    out, tp = inverter.model.forward([mel], noise=[noise], precision="f16f8", taps=["F0", "phase"])
    whole, f0_dev, phase_dev = out[0], tp["F0"][0].reshape(-1), tp["phase"][0].reshape(-1)
    assert whole.shape == (T * hop,) and np.isfinite(whole).all()
    parts, info = inverter.synth_long_from_mel(mel, noise=noise, chunk_frames=400, return_info=True)
    assert info["n_windows"] == 120
    assert np.array_equal(parts, whole), float(np.abs(parts - whole).max())
    print(f"10-minute mel: first chunk after {1e3 * info['first_chunk_latency_s']:.2f} ms, total {1e3 * info['total_s']:.1f} ms "
          f"= {info['audio_s'] / info['total_s']:.0f} audio-s/s (first call: includes allocations)")
    from mbexwn_vocoder_b200.long_form import synth_long
    host, _ = synth_long(inverter.model, mel, noise, 400, "f16f8", 32768, host_loop=True)
    assert np.array_equal(host, whole)
    # the oracle's chunked cumulative phase (tf_wavetable.py:429-492 restated) over the whole ten minutes
    oracle = OracleMBExWN(hp, w, torch.float32)
    v = (f0_dev / np.float32(plan.pulse_rate)).astype(np.float32)[None]
    assert np.array_equal(oracle.stable_cumsum_and_wrap(v)[0], phase_dev)
    # a window 8 s into the signal against the oracle run on the first 980 frames (device F0, same noise)
    n, a, b = 980, 760, 920
    ref = oracle.forward(mel[None, :n], noise[None, :n * spf, None], f0_override=f0_dev[None, :n * ppf])["waveform"][0]
    r, g = ref[a * hop:b * hop].astype(np.float64), whole[a * hop:b * hop].astype(np.float64)
    assert np.abs(g - r).max() <= 1e-4 * np.abs(r).max()
    assert 10 * np.log10(np.sum(r * r) / np.sum((g - r) ** 2)) >= 60.0
