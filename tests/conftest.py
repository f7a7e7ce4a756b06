import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


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
