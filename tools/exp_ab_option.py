#!/usr/bin/env python
"""A/B timing of one engine option on the config-2 step (device-resident): python tools/exp_ab_option.py <option> [reps]."""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter

opt = sys.argv[1]
VALUES = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (0, 1)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0)
eng.set_option("tc_fused", 1)
mels, noise = bench.synthetic_batch(64, 400, plan.steps_per_frame)
pb = eng.prepare([400] * 64, precision="f16f8", with_noise=True)
pb.load(mels, noise)
pb.upload()


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        pb.run_device()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        pb.run_device()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(reps):
    for v in VALUES:
        eng.set_option(opt, v)
        ms = run(20)
        eng.set_option("stage_timing", 1)
        pb.run_device()
        t = pb.wavenet_launch_ms()
        eng.set_option("stage_timing", 0)
        print(f"{opt}={v}: {ms:.3f} ms/step  gate {t['gate']:.3f}  resskip {t['resskip']:.3f}")
