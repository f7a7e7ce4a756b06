import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the model directories of this package hold synthetic configurations without weights: tests run on random-initialised
# weights (MELInverter refuses that unless asked to, like the reference's load_weights would)
os.environ.setdefault("MBEXWN_SYNTHETIC_WEIGHTS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _gpu_ready():
    """A CUDA device and the in-tree library: GPU-marked tests are skipped (not failed) on a CPU-only host."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as e:                      # pragma: no cover
        return False, f"torch unavailable: {e}"
    lib = os.path.join(ROOT, "mbexwn_vocoder_b200", "libmbexwn_b200.so")
    if not os.path.exists(lib):
        return False, "libmbexwn_b200.so not built (python __graft_entry__.py)"
    return True, ""


def pytest_collection_modifyitems(config, items):
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f"needs a B200: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def speech_setup():
    """(hparams, plan, weights) of the synthetic MW-SP-FD model."""
    from mbexwn_vocoder_b200 import get_config_file
    from mbexwn_vocoder_b200 import weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    hp = read_config(get_config_file("SPEECH"))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp["synthetic_weights"]["seed"]))
    return hp, plan, w
