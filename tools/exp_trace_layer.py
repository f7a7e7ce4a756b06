#!/usr/bin/env python
"""Per-tile cycle stamps of the fused WaveNet layer kernel (option "tc_trace") on the config-2 step.

    python tools/exp_trace_layer.py [layer] [out.npy] [precision] [workload]

Prints, for a few CTAs, where the MMA issuer, the TMA producer and one epilogue warp spend their cycles: the share of
the MMA warp's time waiting for operands (`full`) and for a free accumulator (`tmem_empty`), the epilogue's wait for
accumulators (`tmem_full`) and staging tiles.  The kernel is a TRACE instantiation (extra clock reads), so its absolute
time is not a bench number."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mbexwn_vocoder_b200.mel_inverter import MELInverter

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = sys.argv[2] if len(sys.argv) > 2 else None
precision = sys.argv[3] if len(sys.argv) > 3 else "f16f8"
workload = sys.argv[4] if len(sys.argv) > 4 else "config2"
debug = int(sys.argv[5]) if len(sys.argv) > 5 else 0
SLOTS = 384

model_id, batch, frames, _ = bench.WORKLOADS[workload]
inv = MELInverter(model_id, device=0, precision=precision, allow_synthetic_weights=True)
eng, plan = inv.model, inv.plan
eng.set_option("debug_taps", 0)
eng.set_option("tc_cta_group", 2)
eng.set_option("tc_fused", 1)
mels, noise = bench.synthetic_batch(batch, frames, plan.steps_per_frame)
pb = eng.prepare([frames] * batch, precision=precision, with_noise=True)
pb.load(mels, noise)
pb.upload()
for _ in range(3):
    pb.run_device()
torch.cuda.synchronize()
eng.set_option("tc_trace", layer + 1)
eng.set_option("tc_debug", debug)
pb.run_device()
torch.cuda.synchronize()
eng.set_option("tc_trace", 0)
eng.set_option("tc_debug", 0)
print(f"=== layer {layer} {precision} {workload} debug {debug}")
n_cta = torch.cuda.get_device_properties(0).multi_processor_count
buf = np.zeros(n_cta * 3 * SLOTS * 4, dtype=np.uint32)
import ctypes as C
n = eng.lib.mbexwn_tc_trace_read(eng._handle, buf.ctypes.data_as(C.c_void_p), buf.size)
assert n == buf.size, n
tr = buf.reshape(n_cta, 3, SLOTS, 4).astype(np.int64)
if out:
    np.save(out, tr.astype(np.uint32))


def d(a, b):
    return (a - b) & 0xFFFFFFFF


for cta in ((0, 74) if debug else (0, 2, 74, 146)):
    mma, epi, prod = tr[cta, 1], tr[cta, 2], tr[cta, 0]
    nt = int(np.count_nonzero(mma[:, 2]))
    if nt == 0:
        continue
    t0, t1, t2 = mma[:nt, 0], mma[:nt, 1], mma[:nt, 2]
    wfa, wfb = (mma[:nt, 3] >> 16) * 8, (mma[:nt, 3] & 0xFFFF) * 8      # cycles the MMA warp waited for A slabs / B tiles
    wf = wfa + wfb
    total = d(t2[-1], t0[0])
    w_empty = d(t1, t0).sum()
    print(f"cta {cta}: {nt} tiles, MMA warp span {total} cyc; waiting tmem_empty {100 * w_empty / total:.1f} %, "
          f"waiting operands {100 * wf.sum() / total:.1f} % (A {100 * wfa.sum() / total:.1f}, B {100 * wfb.sum() / total:.1f}), issuing {100 * (total - w_empty - wf.sum()) / total:.1f} %")
    ne = int(np.count_nonzero(epi[:, 2]))
    e0, e1, e2, ws = epi[:ne, 0], epi[:ne, 1], epi[:ne, 2], epi[:ne, 3] & 0x7FFFFFFF
    kind = epi[:ne, 3] >> 31
    etot = d(e2[-1], e0[0])
    print(f"   epilogue warp 4: span {etot}; waiting tmem_full {100 * d(e1, e0).sum() / etot:.1f} %, staging {100 * ws.sum() / etot:.1f} %, "
          f"busy per gate tile {np.mean((d(e2, e1) - ws)[kind == 0]):.0f} cyc, per res tile {np.mean((d(e2, e1) - ws)[kind == 1]) if (kind == 1).any() else 0:.0f} cyc")
    npd = int(np.count_nonzero(prod[:, 0]))
    print(f"   producer: waiting for free stages {100 * prod[:npd, 1].sum() / total:.1f} % of the MMA span")
    per = min(nt, 10 if not debug else 0)
    print("   first tiles (mma: wait_empty, wait_full, busy | epi: wait_full, staging, busy):")
    for i in range(per):
        print(f"     tile {i}: mma {d(t1[i], t0[i]):6d} {wf[i]:6d} {d(t2[i], t1[i]) - wf[i]:6d} | epi {d(e1[i], e0[i]):6d} {ws[i]:6d} {d(e2[i], e1[i]) - ws[i]:6d} kind {kind[i]}")
