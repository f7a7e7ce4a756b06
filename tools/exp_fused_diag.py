#!/usr/bin/env python
"""Diagnostic: fused layer kernel variants against the two-launch path on one batch (f16f8): error of wn_out per variant."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbexwn_vocoder_b200.mel_inverter import MELInverter
from oracle.forward import synthetic_mel, synthetic_noise

precision = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
inv = MELInverter("SPEECH", device=0, precision=precision, allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
lengths = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [400, 150, 1, 400, 37, 400, 400, 260]
mels = [synthetic_mel(t, 60 + i) for i, t in enumerate(lengths)]
noise = [synthetic_noise(t * plan.steps_per_frame, 60 + i) for i, t in enumerate(lengths)]
eng.set_option("tc_cta_group", 2)
eng.set_option("tc_fused", 0)
ref, rtp = eng.forward(mels, noise=noise, precision=precision, taps=["wn_out"])
for slab in (0, 1):
    eng.set_option("tc_fused", 1)
    eng.set_option("tc_slab", slab)
    out, tp = eng.forward(mels, noise=noise, precision=precision, taps=["wn_out"])
    errs = []
    for u in range(len(lengths)):
        a, b = rtp["wn_out"][u], tp["wn_out"][u]
        errs.append(float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-30)))
    print(f"{precision} slab={slab}: max |fused - two-launch| / peak per utterance:", " ".join(f"{e:.2e}" for e in errs), flush=True)
    if max(errs) > 1e-3:
        for u in range(len(lengths)):
            a, b = rtp["wn_out"][u].reshape(-1, 30), tp["wn_out"][u].reshape(-1, 30)
            d = np.abs(a - b).max(axis=1)
            bad = np.nonzero(d > 1e-3 * np.abs(a).max())[0]
            if bad.size:
                # runs of bad rows
                runs = np.split(bad, np.nonzero(np.diff(bad) > 1)[0] + 1)
                print(f"  utt {u} ({lengths[u]} frames, {a.shape[0]} rows): {bad.size} bad rows in {len(runs)} runs:",
                      [(int(r[0]), int(r[-1])) for r in runs[:12]])
