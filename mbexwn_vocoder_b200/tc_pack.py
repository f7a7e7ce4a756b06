"""Host-side packing of the WaveNet weights for the tcgen05 tap-GEMM (layout: include/mbexwn.h).

* channels are padded to cpad = ceil(C / 64) * 64 (zero weights, zero bias);
* W1 rows (GEMM N) are permuted so that every 256-row tile (the last may be narrower) is [tanh channels | the matching sigmoid
  channels]: the tanh*sigmoid gate (custom_AE_layers.py:309-321) becomes local to one 128-column accumulator tile;
* res_skip rows are [res channels (cpad) | skip channels (cpad)] (skip only for the last layer);
* every matrix is stored K-major as bf16 [hi | lo] planes with hi + lo ~ the fp32 value, so the same kernel runs
  plain bf16 (hi*hi) or the 3-product split (hi*hi + lo*hi + hi*lo) by listing more K blocks.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import weights as W
from .plan import ModelPlan

TILE_K = 64
GATE_TILE = 256


def hilo(x: np.ndarray) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.cat((hi, lo), dim=1)


def gate_permutation(c: int, cpad: int):
    """For packed row n of W1: (source column in the reference's [tanh(C) | sigmoid(C)] order, valid mask)."""
    n = np.arange(2 * cpad)
    tile = n // GATE_TILE
    width = np.minimum(GATE_TILE, 2 * cpad - tile * GATE_TILE)      # the last tile may be narrower
    within = n - tile * GATE_TILE
    is_sig = within >= width // 2
    ch = (GATE_TILE // 2) * tile + np.where(is_sig, within - width // 2, within)
    return np.where(is_sig, c + ch, ch), ch < c


def pack_tc_weights(plan: ModelPlan, weights: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    wn = plan.wavenet
    C, k = wn.c, wn.k
    cpad = -(-C // TILE_K) * TILE_K
    name = wn.name + "_WNBlock_WN"
    col1, ok1 = gate_permutation(C, cpad)
    out: Dict[str, torch.Tensor] = {}
    for i in range(wn.n_layers):
        w, b = W.folded(weights, f"{name}/conv1D_{i}")                # (k, C, 2C), (2C,)
        w1 = np.zeros((2 * cpad, k, cpad), dtype=np.float32)
        w1[ok1, :, :C] = np.transpose(w[:, :, col1[ok1]], (2, 0, 1))
        b1 = np.zeros(2 * cpad, dtype=np.float32)
        b1[ok1] = b[col1[ok1]]
        out[f"{name}/tc/W1_{i}"] = hilo(w1.reshape(2 * cpad, k * cpad))
        out[f"{name}/tc/b1_{i}"] = torch.from_numpy(b1)
        r, rb = W.folded(weights, f"{name}/res_skip_{i}")             # (1, C, 2C) or (1, C, C) for the last layer
        last = i == wn.n_layers - 1
        n2 = cpad if last else 2 * cpad
        m = np.arange(n2)
        if last:
            chan, col2 = m, m
        else:
            chan = np.where(m < cpad, m, m - cpad)
            col2 = np.where(m < cpad, m, C + (m - cpad))
        ok2 = chan < C
        r2 = np.zeros((n2, cpad), dtype=np.float32)
        r2[ok2, :C] = r[0][:, col2[ok2]].T
        rb2 = np.zeros(n2, dtype=np.float32)
        rb2[ok2] = rb[col2[ok2]]
        out[f"{name}/tc/R_{i}"] = hilo(r2)
        out[f"{name}/tc/rb_{i}"] = torch.from_numpy(rb2)
    return out
