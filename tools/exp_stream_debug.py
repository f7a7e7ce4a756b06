#!/usr/bin/env python
"""Which of {blocking run_host, CUDA-graph replay, pipelined begin/wait} disagree on small batches?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbexwn_vocoder_b200.mel_inverter import MELInverter
from oracle.forward import synthetic_mel, synthetic_noise

inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
plan, eng = inv.plan, inv.model
batches, noises = [], []
for b in range(5):
    lengths = [12 + 3 * b, 7 + b] if b % 2 else [20, 9, 5]
    batches.append([synthetic_mel(t, 10 * b + i) for i, t in enumerate(lengths)])
    noises.append([synthetic_noise(t * plan.steps_per_frame, 10 * b + i) for i, t in enumerate(lengths)])
eng.use_graphs = False
plain = [inv.synth_batch(m, noise=z) for m, z in zip(batches, noises)]
plain2 = [inv.synth_batch(m, noise=z) for m, z in zip(batches, noises)]
eng.use_graphs = True
graph = [inv.synth_batch(m, noise=z) for m, z in zip(batches, noises)]
graph2 = [inv.synth_batch(m, noise=z) for m, z in zip(batches, noises)]
stream = list(inv.synth_stream(batches, noise=noises))
stream2 = list(inv.synth_stream(batches, noise=noises))


def same(a, b):
    return ["".join("=" if np.array_equal(x, y) else "X" for x, y in zip(p, q)) for p, q in zip(a, b)]


print("plain  vs plain2 ", same(plain, plain2))
print("plain  vs graph  ", same(plain, graph))
print("plain  vs graph2 ", same(plain, graph2))
print("plain  vs stream ", same(plain, stream))
print("plain  vs stream2", same(plain, stream2))
