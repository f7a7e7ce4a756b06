"""Weight container for the MBExWN forward path.

The reference stores, per weight-normalised conv, a direction tensor ``v`` (k, cin, cout), a gain ``g``
(cout,) and a ``bias`` (cout,) and re-normalises on every call (conv_layers.py:149-154); PReLU layers hold
``alpha`` per channel (shared over time, custom_pulsed_generator.py:247-250).  This module

* creates Keras-like random-initialised weights for a plan (the released checkpoints are not available
  offline; SURVEY.md 8d "Synthetic inputs"),
* folds weight-norm once at load (``fold``), and
* reads/writes the un-folded container as ``.npz`` (names ``<layer>/v|g|bias``, ``<act>/alpha``).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .plan import ACT_PRELU, ConvLayer, ModelPlan


def _glorot_uniform(rng, k, cin, cout):
    limit = np.sqrt(6.0 / (k * cin + k * cout))
    return rng.uniform(-limit, limit, size=(k, cin, cout))


def _init_conv(rng, layer: ConvLayer, gain_jitter: float, bias_std: float) -> Dict[str, np.ndarray]:
    if layer.init_std is not None:
        v = rng.normal(0.0, layer.init_std, size=(layer.k, layer.cin, layer.cout))
    else:
        v = _glorot_uniform(rng, layer.k, layer.cin, layer.cout)
    if layer.cb_free:
        # checkerboard-free sub-pixel init: kernels identical across the unfold factor (conv_layers.py:73-77)
        f = layer.cb_free
        v = v.reshape(layer.k, layer.cin, f, layer.cout // f).mean(axis=-2, keepdims=True)
        v = np.tile(v, (1, 1, f, 1)).reshape(layer.k, layer.cin, layer.cout)
    g = np.linalg.norm(v.reshape(-1, layer.cout), axis=0)          # conv_layers.py:86
    if gain_jitter:
        g = g * rng.uniform(1.0 - gain_jitter, 1.0 + gain_jitter, size=g.shape)
    bias = rng.normal(0.0, bias_std, size=(layer.cout,)) if bias_std else np.zeros(layer.cout)
    return {f"{layer.name}/v": v.astype(np.float32), f"{layer.name}/g": g.astype(np.float32),
            f"{layer.name}/bias": bias.astype(np.float32)}


def init_synthetic(plan: ModelPlan, seed: int = 0, lively: bool = True) -> Dict[str, np.ndarray]:
    """Random weights with the reference's initialisers.

    ``lively=False`` is the plain Keras state at construction (zero biases, g = ||v||, alpha = plan.alpha).
    ``lively=True`` additionally jitters gains/biases/alphas and boosts the two sub-net heads so that the
    F0 contour sweeps its range and the cepstral envelope is non-trivial -- a freshly initialised net is a
    near-constant F0 and a flat filter, which would leave most of the excitation / VTF arithmetic untested.
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for layer in plan.conv_layers():
        out.update(_init_conv(rng, layer, 0.25 if lively else 0.0, 0.05 if lively else 0.0))
    for ops in (plan.pp_ops, plan.ps_ops):
        for op in ops:
            if op.act == ACT_PRELU and op.act_name:
                a = np.full(op.act_channels, plan.alpha)
                if lively:
                    a = a * rng.uniform(0.5, 1.5, size=a.shape)
                out[f"{op.act_name}/alpha"] = a.astype(np.float32)
    if lively:
        # sub-net trunks: N(0, 0.02) init leaves activations tiny; scale gains so features are O(1)
        for ops, head_gain in ((plan.pp_ops, 25.0), (plan.ps_ops, 1.5)):
            convs = [op.conv for op in ops if op.kind == "conv"]
            if not convs:                                      # ps_off: no PS sub-net
                continue
            for layer in convs[:-1]:
                out[f"{layer.name}/g"] = (out[f"{layer.name}/g"] * 3.0).astype(np.float32)
            out[f"{convs[-1].name}/g"] = (out[f"{convs[-1].name}/g"] * head_gain).astype(np.float32)
    return out


def fold(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """W = g * v / max(||v||_2 over (k, cin), 1e-6)  -- tf.nn.l2_normalize(axis=[0, 1]), epsilon 1e-12 on the square."""
    v64 = v.astype(np.float64)
    sq = np.sum(v64 * v64, axis=(0, 1), keepdims=True)
    return (g.astype(np.float64) * v64 / np.sqrt(np.maximum(sq, 1e-12))).astype(np.float32)


def folded(weights: Dict[str, np.ndarray], name: str):
    """(W (k, cin, cout) float32, bias (cout,) float32) of one conv layer."""
    return fold(weights[f"{name}/v"], weights[f"{name}/g"]), weights[f"{name}/bias"].astype(np.float32)


def check(plan: ModelPlan, weights: Dict[str, np.ndarray]) -> None:
    """Raise KeyError/ValueError if a container does not match the plan's layers."""
    for layer in plan.conv_layers():
        v = weights[f"{layer.name}/v"]
        if v.shape != (layer.k, layer.cin, layer.cout):
            raise ValueError(f"{layer.name}/v has shape {v.shape}, expected {(layer.k, layer.cin, layer.cout)}")
        for part in ("g", "bias"):
            if weights[f"{layer.name}/{part}"].shape != (layer.cout,):
                raise ValueError(f"{layer.name}/{part} has the wrong shape")
    for ops in (plan.pp_ops, plan.ps_ops):
        for op in ops:
            if op.act == ACT_PRELU and op.act_name and weights[f"{op.act_name}/alpha"].shape != (op.act_channels,):
                raise ValueError(f"{op.act_name}/alpha has the wrong shape")


def save(path: str, weights: Dict[str, np.ndarray]) -> None:
    np.savez(path, **weights)


def load(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as data:
        return {k: data[k] for k in data.files}
