#!/usr/bin/env python
"""Mint seeds-to-tensors golden vectors from the CPU oracle (float32) for the synthetic MW-SP-FD model.

The reference ships no golden vectors, so these pin the oracle against accidental change and give
the GPU tests a fixture that does not depend on running the oracle.  Output: tests/golden/oracle_speech_T16.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mbexwn_vocoder_b200 import get_config_file, weights as W  # noqa: E402
from mbexwn_vocoder_b200.config import read_config  # noqa: E402
from mbexwn_vocoder_b200.plan import build_plan  # noqa: E402
from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise  # noqa: E402


def main():
    torch.set_num_threads(1)          # fixed summation order inside torch's conv kernels
    hp = read_config(get_config_file("SPEECH"))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp["synthetic_weights"]["seed"]))
    orc = OracleMBExWN(hp, w, torch.float32)
    T = 16
    mel, noise = synthetic_mel(T, 0), synthetic_noise(T * plan.steps_per_frame, 0)
    r = orc.forward(mel[None], noise[None])
    out = {"mel": mel, "noise": noise}
    for k in ("F0", "phase", "index", "pulse", "subbands", "excitation", "ceps", "waveform"):
        out[k] = np.asarray(r[k][0])
    out["vtf_abs"] = np.abs(np.asarray(r["vtf"][0])).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "oracle_speech_T16.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
