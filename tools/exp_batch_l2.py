#!/usr/bin/env python
"""Does a sub-batch whose residual stream + gated activations fit the 126 MB L2 run the WaveNet faster per row?
Device-resident timing of the config-2 step at several batch sizes (5 s utterances): python tools/exp_batch_l2.py [sizes]."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 32, 16, 10, 8, 6, 5, 4, 3, 64]
inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0)
for B in sizes:
    mels, noise = bench.synthetic_batch(B, 400, plan.steps_per_frame)
    pb = eng.prepare([400] * B, precision="f16f8", with_noise=True)
    pb.load(mels, noise)
    pb.upload()
    n = max(20, 1280 // B)
    for _ in range(5):
        pb.run_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        pb.run_device()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.set_option("stage_timing", 1)
    pb.run_device()
    t = pb.wavenet_launch_ms()
    st = pb.stage_ms()
    eng.set_option("stage_timing", 0)
    print(f"B={B:3d}: {ms:7.3f} ms/step = {B * 5 / ms * 1e3:8.0f} audio-s/s | per utterance: step {ms / B * 1e3:6.1f} us, gate "
          f"{t['gate'] / B * 1e3:6.1f} us, resskip {t['resskip'] / B * 1e3:6.1f} us, wavenet {st['wavenet'] / B * 1e3:6.1f} us", flush=True)
