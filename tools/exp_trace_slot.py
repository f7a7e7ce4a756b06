#!/usr/bin/env python
"""Turn-around of ONE operand ring slot of the fused layer kernel (tc_debug 32): for B slot 0 the producer lane logs when it saw
the slot empty and when its TMA was out, the MMA warp when it saw the slot full and when its commit was issued."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter
import ctypes as C
extra = int(sys.argv[1]) if len(sys.argv) > 1 else 0
inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0); eng.set_option("tc_cta_group", 2); eng.set_option("tc_fused", 1)
mels, noise = bench.synthetic_batch(64, 400, plan.steps_per_frame)
pb = eng.prepare([400] * 64, precision="f16f8", with_noise=True)
pb.load(mels, noise); pb.upload()
for _ in range(3):
    pb.run_device()
torch.cuda.synchronize()
eng.set_option("tc_trace", 4); eng.set_option("tc_debug", 32 | extra)
pb.run_device(); torch.cuda.synchronize()
eng.set_option("tc_trace", 0); eng.set_option("tc_debug", 0)
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
buf = np.zeros(n_cta * 3 * 384 * 4, dtype=np.uint32)
assert eng.lib.mbexwn_tc_trace_read(eng._handle, buf.ctypes.data_as(C.c_void_p), buf.size) == buf.size
tr = buf.reshape(n_cta, 3, 384, 4).astype(np.int64)
d = lambda a, b: (a - b) & 0xFFFFFFFF
for cta in (0, 74):
    prod, mma = tr[cta, 0], tr[cta, 1]
    n = min(int(np.count_nonzero(prod[:, 1])), int(np.count_nonzero(mma[:, 1]))) - 1
    seen, out = prod[:n, 0], prod[:n, 1]
    full, commit = mma[:n, 0], mma[:n, 1]
    print(f"cta {cta} (tc_debug {32 | extra}), B slot 0, fills 8..{n}: median cycles")
    k = slice(8, n)
    print("   commit issued (fill f-1) -> empty seen by the producer lane :", int(np.median(d(seen[1:n], commit[:n - 1])[7:])))
    print("   empty seen -> TMA issued                                    :", int(np.median(d(out, seen)[k])))
    print("   TMA issued -> full seen by the MMA warp                     :", int(np.median(d(full, out)[k])))
    print("   full seen -> commit issued                                  :", int(np.median(d(commit, full)[k])))
    print("   whole turn-around (commit to commit)                        :", int(np.median(d(commit[1:n], commit[:n - 1])[7:])))
