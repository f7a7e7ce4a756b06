#!/usr/bin/env python
"""Tiny end-to-end runs for compute-sanitizer (memcheck / racecheck): every precision, a ragged batch, the analysis kernel."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbexwn_vocoder_b200.mel_inverter import MELInverter
from oracle.forward import synthetic_mel
from oracle.analysis import synthetic_audio

inv = MELInverter("SPEECH", device=0, precision="f16f8")
mels = [synthetic_mel(t, i) for i, t in enumerate((9, 1, 14))]
for prec in ("f16f8", "bf16x3", "bf16", "fp32"):
    inv.precision = prec
    out = inv.synth_batch(mels)
    assert all(np.isfinite(w).all() for w in out), prec
    print(prec, "ok", [w.shape for w in out])
dd = inv.generate_mel_from_snd(synthetic_audio(2000, 0), 24000)
print("analysis ok", dd["mell"].shape)
long = inv.synth_long_from_mel(synthetic_mel(45, 3), chunk_frames=20)
print("long ok", long.shape)
