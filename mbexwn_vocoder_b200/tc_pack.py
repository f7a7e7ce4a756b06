"""Host-side packing of the WaveNet weights for the tcgen05 tap-GEMM (layout: include/mbexwn.h).

* channels are padded to cpad = ceil(C / 64) * 64 (zero weights, zero bias);
* W1 rows (GEMM N) are permuted so that every 256-row tile (the last may be narrower) is [tanh channels | the matching sigmoid
  channels]: the tanh*sigmoid gate (custom_AE_layers.py:309-321) becomes local to one 128-column accumulator tile;
* res_skip rows are [res channels (cpad) | WaveNet output channels (c_out padded to 32)] (output only for the last layer):
  the skip sum feeds nothing but the linear `end` 1x1 (custom_AE_layers.py:337-340), so the skip half of every res_skip
  matrix is pre-multiplied by W_end in float64 (R_skip @ W_end, C x c_out) and the kernel accumulates end(skip) directly;
  all skip biases and the `end` bias travel with layer 0;
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
OUT_PAD = 32


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


GATE_CHUNK = 16


def gate_permutation_chunked(c: int, cpad: int):
    """Row order of W1 for the fused layer kernel (csrc/k_wavenet_layer.cu): per 16-channel chunk [16 tanh rows | the 16 matching
    sigmoid rows], so that any N tile made of whole chunks holds both halves of the gate and a chunk is 32 adjacent
    accumulator columns.  Returns (source column in the reference's [tanh(C) | sigmoid(C)] order, valid mask)."""
    n = np.arange(2 * cpad)
    within = n % (2 * GATE_CHUNK)
    is_sig = within >= GATE_CHUNK
    ch = GATE_CHUNK * (n // (2 * GATE_CHUNK)) + within % GATE_CHUNK
    return np.where(is_sig, c + ch, ch), ch < c


def pack_tc_weights(plan: ModelPlan, weights: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    """bf16 [hi | lo] planes of the packed matrices, fp32 biases."""
    out: Dict[str, torch.Tensor] = {}
    for key, m in _tc_matrices(plan, weights).items():
        out[key] = hilo(m) if m.ndim == 2 else torch.from_numpy(m)
    return out


def _tc_matrices(plan: ModelPlan, weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """fp32 GEMM operands of the tensor-core path for every WaveNet block (one for the released models)."""
    out: Dict[str, np.ndarray] = {}
    for wn in plan.blocks:
        out.update(_tc_block_matrices(wn, weights))
    return out


def _tc_block_matrices(wn, weights: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """One WaveNetAE: W1_i (2 cpad, k cpad), R_i (n2, cpad) and the biases b1_i, rb_i."""
    C, k = wn.c, wn.k
    cpad = -(-C // TILE_K) * TILE_K
    name = wn.name + "_WNBlock_WN"
    col1, ok1 = gate_permutation(C, cpad)
    col1f, ok1f = gate_permutation_chunked(C, cpad)
    out: Dict[str, np.ndarray] = {}
    we, be = W.folded(weights, f"{name}/end")                         # (1, C, c_out), (c_out,)
    we64 = we[0].astype(np.float64)
    c_out = we64.shape[1]
    out_pad = -(-c_out // OUT_PAD) * OUT_PAD
    skip_bias = np.zeros(C, dtype=np.float64)
    rbs = []
    for i in range(wn.n_layers):
        w, b = W.folded(weights, f"{name}/conv1D_{i}")                # (k, C, 2C), (2C,)
        w1 = np.zeros((2 * cpad, k, cpad), dtype=np.float32)
        w1[ok1, :, :C] = np.transpose(w[:, :, col1[ok1]], (2, 0, 1))
        b1 = np.zeros(2 * cpad, dtype=np.float32)
        b1[ok1] = b[col1[ok1]]
        out[f"{name}/tc/W1_{i}"] = w1.reshape(2 * cpad, k * cpad)
        out[f"{name}/tc/b1_{i}"] = b1
        w1f = np.zeros((2 * cpad, k, cpad), dtype=np.float32)
        w1f[ok1f, :, :C] = np.transpose(w[:, :, col1f[ok1f]], (2, 0, 1))
        b1f = np.zeros(2 * cpad, dtype=np.float32)
        b1f[ok1f] = b[col1f[ok1f]]
        out[f"{name}/tcf/W1_{i}"] = w1f.reshape(2 * cpad, k * cpad)
        out[f"{name}/tcf/b1_{i}"] = b1f
        r, rb = W.folded(weights, f"{name}/res_skip_{i}")             # (1, C, 2C) or (1, C, C) for the last layer
        last = i == wn.n_layers - 1
        r64 = r[0].astype(np.float64)
        r_skip = r64 if last else r64[:, C:]
        n_res = 0 if last else cpad
        n2 = n_res + out_pad
        r2 = np.zeros((n2, cpad), dtype=np.float32)
        rb2 = np.zeros(n2, dtype=np.float32)
        if not last:
            r2[:C, :C] = r[0][:, :C].T
            rb2[:C] = rb[:C]
        r2[n_res:n_res + c_out, :C] = (r_skip @ we64).T.astype(np.float32)
        skip_bias += (rb if last else rb[C:]).astype(np.float64)
        out[f"{name}/tc/R_{i}"] = r2
        rbs.append((f"{name}/tc/rb_{i}", rb2, n_res))
    # wn_out = sum_i act_i @ (Rskip_i @ We) + (sum_i bskip_i) @ We + be: the constant term rides on layer 0 (assign)
    key0, rb0, n_res0 = rbs[0]
    rb0[n_res0:n_res0 + c_out] = (skip_bias @ we64 + be.astype(np.float64)).astype(np.float32)
    for key, rb2, _ in rbs:
        out[key] = rb2
    return out


# ---- MBEXWN_PREC_F16F8: fp16 main product + two e4m3 correction products ---------------------------------------
# x * w ~ f16(x) f16(w) + 2^-15 [ e4m3((x - f16 x) 2^sa) e4m3(w 2^(15 - sa)) + e4m3(x 2^sa') e4m3((w - f16 w) 2^(15 - sa')) ]
# The tensor cores rescale the accumulated correction products by 2^-15 (scale-input-d of the first fp16 MMA of a tile), so
# for every GEMM the plane scales of the two operands must multiply to 2^15.  The activation hi8 planes are unscaled
# (e4m3(x): |x| up to 448, so the weight lo8 planes carry 2^15); the activation lo8 shifts are kernel options (tc8_h_lo for
# the residual stream, tc8_a_lo for the gated activations in (-1, 1)):
CORR_SHIFT = 15
TC8_DEFAULT_SHIFTS = {"tc8_h_lo": 9, "tc8_a_lo": 10}
E4M3_MAX = 448.0


def f16f8_planes(x: np.ndarray, hi_shift: int, lo_shift: int) -> torch.Tensor:
    """(n, K) fp32, K % 64 == 0 -> (n, 4 K) uint8 rows [fp16 (K) | per 64 of K: e4m3(x 2^hi_shift) (64), e4m3((x - f16 x)
    2^lo_shift) (64)]: the weight side of the layout in include/mbexwn.h (activations store lo8 first)."""
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    n, k = t.shape
    h16 = t.to(torch.float16)
    lo = t - h16.to(torch.float32)
    hi8 = torch.clamp(t * 2.0 ** hi_shift, -E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)
    lo8 = torch.clamp(lo * 2.0 ** lo_shift, -E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)
    f8 = torch.stack((hi8.reshape(n, k // TILE_K, TILE_K), lo8.reshape(n, k // TILE_K, TILE_K)), dim=2).reshape(n, 2 * k)
    return torch.cat((h16.view(torch.uint8), f8), dim=1).contiguous()


def choose_tc8_shifts(w1_max: float, r_max: float) -> Dict[str, int]:
    """Activation-plane shifts; the weight hi8 plane 2^(15 - *_lo) must not saturate e4m3 for the largest weight."""
    out = dict(TC8_DEFAULT_SHIFTS)
    for key, wmax in (("tc8_h_lo", w1_max), ("tc8_a_lo", r_max)):
        room = int(np.floor(np.log2(E4M3_MAX / max(wmax, 1e-30))))          # largest weight-side shift without saturation
        out[key] = max(out[key], CORR_SHIFT - room)
    return out


def pack_tc8_weights(plan: ModelPlan, weights: Dict[str, np.ndarray]):
    """The matrices of pack_tc_weights (same row order / folding, fp32 before the split) as [fp16 | e4m3 | e4m3] planes.

    Returns ({name: uint8 tensor}, shifts)."""
    mats = _tc_matrices(plan, weights)
    w1_max = max(float(np.abs(m).max()) for k, m in mats.items() if "/W1_" in k)
    r_max = max(float(np.abs(m).max()) for k, m in mats.items() if "/R_" in k)
    sh = choose_tc8_shifts(w1_max, r_max)
    out: Dict[str, torch.Tensor] = {}
    for key, m in mats.items():
        if "/W1_" in key:
            out[key.replace("/tcf/", "/tcf8/").replace("/tc/", "/tc8/")] = f16f8_planes(m, CORR_SHIFT - sh["tc8_h_lo"], CORR_SHIFT)
        elif "/R_" in key:
            out[key.replace("/tc/", "/tc8/")] = f16f8_planes(m, CORR_SHIFT - sh["tc8_a_lo"], CORR_SHIFT)
    return out, sh


def pack_conv_tc(w: np.ndarray) -> torch.Tensor:
    """(k, cin, cout) folded conv kernel -> (cout, [hi | lo] x k x cin_pad) bf16, K-major, cin padded to 64."""
    k, cin, cout = w.shape
    cin_pad = -(-cin // TILE_K) * TILE_K
    b = np.zeros((cout, k, cin_pad), dtype=np.float32)
    b[:, :, :cin] = np.transpose(w, (2, 0, 1))
    return hilo(b.reshape(cout, k * cin_pad))


def pack_subnet_weights(plan: ModelPlan, weights: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    """Tensor-core copies ("<layer>/tc/W") of the wide mel-rate convs: both sub-nets and the conditioning conv."""
    layers = [op.conv for ops in (plan.pp_ops, plan.ps_ops) for op in ops if op.kind == "conv"]
    cond_names = {f"{wn.name}_WNBlock_WN/cond_" for wn in plan.blocks}
    layers += [l for l in plan.conv_layers() if l.name in cond_names]
    out: Dict[str, torch.Tensor] = {}
    for layer in layers:
        if layer.cin >= 32 and layer.cout >= 16 and layer.cout % 8 == 0 and layer.dilation == 1:
            w, _ = W.folded(weights, layer.name)
            out[f"{layer.name}/tc/W"] = pack_conv_tc(w)
    return out
