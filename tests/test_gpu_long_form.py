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
