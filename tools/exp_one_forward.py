#!/usr/bin/env python
"""One config-2 forward (device-resident) with engine options from the command line: python tools/exp_one_forward.py opt=value ...
(to run under ncu: ncu -k regex:wn_layer -c 3 --metrics ... python tools/exp_one_forward.py tc_cluster=4)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter

inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0)
for a in sys.argv[1:]:
    k, v = a.split("=")
    eng.set_option(k, int(v))
mels, noise = bench.synthetic_batch(64, 400, plan.steps_per_frame)
pb = eng.prepare([400] * 64, precision="f16f8", with_noise=True)
pb.load(mels, noise)
pb.upload()
pb.run_device()
torch.cuda.synchronize()
print("cluster used", eng.get_info("tc_last_cluster"), "max quads", eng.get_info("tc_max_quads"), flush=True)
