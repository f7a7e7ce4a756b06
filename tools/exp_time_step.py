#!/usr/bin/env python
"""Device-resident time of the config-2 step and of its WaveNet launches (A/B runs of two library builds): prints one line."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter

inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0)
for a in sys.argv[2:]:                      # opt=value ... (options a library build does not know are skipped)
    k, v = a.split("=")
    try:
        eng.set_option(k, int(v))
    except Exception:
        pass
mels, noise = bench.synthetic_batch(64, 400, plan.steps_per_frame)
pb = eng.prepare([400] * 64, precision="f16f8", with_noise=True)
pb.load(mels, noise)
pb.upload()
for _ in range(5):
    pb.run_device()
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        pb.run_device()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 30)
eng.set_option("stage_timing", 1)
pb.run_device()
t = pb.wavenet_launch_ms()
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''}: step {min(res):.3f} / {sorted(res)[1]:.3f} ms (min / median of 3 x 30), gate {t['gate']:.3f} resskip {t['resskip']:.3f}", flush=True)
