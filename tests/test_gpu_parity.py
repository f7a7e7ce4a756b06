"""GPU parity: CUDA path (through the C-ABI) vs the CPU oracle on identical weights, mels and noise.

Tolerances are the north-star's: integer wavetable indices bit-exact; fp32 path per-stage max-abs error
<= 1e-4 of the stage peak and waveform SNR >= 60 dB.
"""
import numpy as np
import pytest
import torch

from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise

pytestmark = pytest.mark.gpu

STAGE_TOL = 1e-4


def _rel_err(a, b):
    peak = float(np.abs(b).max())
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max()) / max(peak, 1e-30)


def _snr_db(ref, test):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum(err ** 2), 1e-300))


@pytest.fixture(scope="module")
def engine(speech_setup):
    from mbexwn_vocoder_b200.engine import Engine
    hp, plan, w = speech_setup
    eng = Engine(plan, w, device=0)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def oracle(speech_setup):
    hp, plan, w = speech_setup
    return OracleMBExWN(hp, w, torch.float32)


def _case(lengths, plan):
    mels = [synthetic_mel(t, i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, i) for i, t in enumerate(lengths)]
    return mels, noise


STAGES = ["F0", "pulse", "wn_in", "cond", "skip", "subbands", "excitation", "ceps", "waveform"]


def _oracle_taps(oracle, mel, noise, f0=None):
    r = oracle.forward(mel[None], noise[None], f0_override=None if f0 is None else f0[None])
    r["cond"] = r["cond_lo"]
    return {k: np.asarray(v)[0] for k, v in r.items() if isinstance(v, np.ndarray)}


@pytest.mark.parametrize("lengths", [[40], [23, 57, 10]])
def test_stage_parity_fp32(engine, oracle, speech_setup, lengths):
    hp, plan, w = speech_setup
    mels, noise = _case(lengths, plan)
    taps = ["F0", "phase", "index", "pulse", "wn_in", "cond", "skip", "subbands", "excitation", "ceps"]
    out, tp = engine.forward(mels, noise=noise, precision="fp32", taps=taps)
    for u, t in enumerate(lengths):
        ref = _oracle_taps(oracle, mels[u], noise[u])
        got = {k: v[u] for k, v in tp.items()}
        got["waveform"] = out[u]
        assert out[u].shape == (t * plan.hop,)
        for st in STAGES:
            e = _rel_err(got[st].reshape(-1), ref[st].reshape(-1))
            print(f"utt {u} T={t} stage {st}: max|err|/peak = {e:.3e}")
        # the F0 contour differs in the last bits (summation order), so the phase can flip an index at isolated
        # samples here; the bit-exact index test feeds both sides the same F0 (test_index_bit_exact)
        for st in ["F0", "cond", "ceps"]:
            assert _rel_err(got[st].reshape(-1), ref[st].reshape(-1)) <= STAGE_TOL, st


@pytest.mark.parametrize("lengths", [[40], [23, 57, 10], [101]])
def test_index_bit_exact_and_downstream(engine, oracle, speech_setup, lengths):
    """Same F0 into both sides: phase and integer table index bit-exact, every later stage within 1e-4 of peak."""
    hp, plan, w = speech_setup
    mels, noise = _case(lengths, plan)
    f0 = [oracle.generate_f0(torch.as_tensor(m[None])).numpy()[0] for m in mels]
    taps = ["F0", "phase", "index", "pulse", "wn_in", "cond", "skip", "subbands", "excitation", "ceps"]
    out, tp = engine.forward(mels, noise=noise, f0=f0, precision="fp32", taps=taps)
    for u, t in enumerate(lengths):
        ref = _oracle_taps(oracle, mels[u], noise[u], f0[u])
        assert np.array_equal(tp["index"][u], ref["index"]), "wavetable index not bit-exact"
        assert np.array_equal(tp["phase"][u], ref["phase"]), "wrapped phase not bit-exact"
        got = {k: v[u] for k, v in tp.items()}
        got["waveform"] = out[u]
        for st in STAGES[1:]:
            e = _rel_err(got[st].reshape(-1), ref[st].reshape(-1))
            print(f"utt {u} T={t} stage {st}: max|err|/peak = {e:.3e}")
            assert e <= STAGE_TOL, st
        assert _snr_db(ref["waveform"], out[u]) >= 60.0
