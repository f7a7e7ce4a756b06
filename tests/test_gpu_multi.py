"""The many-utterance front end (multi_gpu.py): LPT shards, batches under a frame budget, pipelined host forward, host gather."""
import numpy as np
import pytest

from oracle.forward import synthetic_mel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def inv():
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    m = MELInverter("SING", device=0, precision="f16f8")
    yield m
    m.model.close()


def test_synth_many_equals_one_batch_and_does_not_depend_on_the_budget(inv):
    """BASELINE.json configs[3] in miniature: a ragged set through synth_many (batches of at most 1500 / 4000 frames, two
    pinned buffer sets, in-kernel noise keyed by the global utterance id) gives bit for bit what ONE forward over the whole
    set gives -- the property that makes sampled utterances of the sharded 8192-set comparable with a single-GPU run."""
    rng = np.random.default_rng(1)
    lengths = [int(t) for t in rng.integers(80, 2401, size=20)]
    lengths[:3] = [80, 2400, 1]
    mels = [synthetic_mel(t, 300 + i) for i, t in enumerate(lengths)]
    whole = inv.model.forward(mels, precision="f16f8", seed=inv.seed, utt_ids=list(range(len(mels))))[0]
    for budget in (1500, 4000, 32768):
        got, stats = inv.synth_many(mels, max_batch_frames=budget, return_stats=True)
        assert len(got) == len(mels) and stats[0].n_utts == len(mels) and stats[0].range_reruns == 0
        for i, (a, b) in enumerate(zip(whole, got)):
            assert a.shape == b.shape == (lengths[i] * inv.plan.hop,)
            assert np.array_equal(a, b), (budget, i)


def test_shards_of_a_pool_reproduce_the_single_device_result(inv):
    """Two engines (here on one B200; MELInverter(devices=[0, 1]) puts them on two) share the set by LPT: the gathered list
    equals the single-engine list bit for bit, every utterance exactly once."""
    from mbexwn_vocoder_b200.engine import Engine
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    from mbexwn_vocoder_b200.multi_gpu import DevicePool, imbalance
    from mbexwn_vocoder_b200 import sched
    rng = np.random.default_rng(2)
    lengths = [int(t) for t in rng.integers(30, 900, size=17)]
    mels = [synthetic_mel(t, 500 + i) for i, t in enumerate(lengths)]
    one = inv.synth_many(mels, max_batch_frames=3000)
    other = MELInverter.__new__(MELInverter)
    other.__dict__.update(inv.__dict__)
    other.model = Engine(inv.plan, inv.weights, device=0)
    other._pool = None
    two, stats = DevicePool([inv, other]).synth_many(mels, max_batch_frames=3000, seed=inv.seed, precision="f16f8", return_stats=True)
    other.model.close()
    assert sum(s.n_utts for s in stats) == len(mels) and all(s.n_utts > 0 for s in stats)
    assert imbalance(lengths, sched.lpt_shards(lengths, 2)) < 1.05
    for a, b in zip(one, two):
        assert np.array_equal(a, b)


def test_range_guard_word_is_clear_on_the_synthetic_models(inv):
    """mbexwn_range_status: the f16f8 kernels flag residual-stream values beyond the e4m3 hi8 plane (|x| > 448); the synthetic
    models stay far below, and an inflated start conv raises it (and MELInverter falls back to bf16x3, test_gpu_parity.py)."""
    inv.model.range_status(reset=True)
    inv.synth_batch([synthetic_mel(40, 1)])
    assert inv.model.range_status(reset=True) == 0
