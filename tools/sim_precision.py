#!/usr/bin/env python
"""CPU study of operand-split schemes for the WaveNet contractions (design aid for csrc/k_wavenet_tc.cu).

Emulates, with exact fp64 accumulation of the *rounded* operands, what the tensor cores would compute:
  bf16x3   : bf16(a)*bf16(b) + bf16(a_lo)*bf16(b) + bf16(a)*bf16(b_lo)                       (3 bf16-rate products)
  f16      : f16(a)*f16(b)                                                                   (1)
  f16f8    : f16(a)*f16(b) + 2^-S [ e4m3(a_lo 2^sa) e4m3(b 2^sb) + e4m3(a 2^sa') e4m3(b_lo 2^sb') ]
             (1 f16-rate product + 2 fp8 products at twice the rate = 2 units)
and reports the error of the WaveNet output against the fp64 run of the oracle.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mbexwn_vocoder_b200 import get_config_file, weights as W      # noqa: E402
from mbexwn_vocoder_b200.config import read_config                # noqa: E402
from mbexwn_vocoder_b200.plan import build_plan                   # noqa: E402
from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise, weight_norm_kernel   # noqa: E402


def rnd(x, dt):
    return x.to(dt).to(torch.float64)


def e4m3(x):
    return x.clamp(-448, 448).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


_E2M1 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)


def fp4_block(x, axis, block, scale_kind):
    """Block-scaled e2m1 emulation along `axis`: nvf4 = e4m3 scale per 16 values, mxf4 = power-of-two scale per 32."""
    x = x.movedim(axis, -1)
    shp = x.shape
    K = shp[-1]
    pad = (-K) % block
    if pad:
        x = torch.cat([x, torch.zeros(*shp[:-1], pad, dtype=x.dtype)], dim=-1)
    xb = x.reshape(*shp[:-1], -1, block)
    amax = xb.abs().amax(dim=-1, keepdim=True).clamp_min(1e-300)
    if scale_kind == "nvf4":
        sc = e4m3_pos(amax / 6.0)
    else:
        sc = 2.0 ** torch.ceil(torch.log2(amax / 6.0))
    q = xb / sc
    mag = q.abs().clamp(max=6.0)
    idx = (mag.unsqueeze(-1) - _E2M1).abs().argmin(dim=-1)
    deq = torch.sign(q) * _E2M1[idx] * sc
    out = deq.reshape(*shp[:-1], -1)[..., :K]
    return out.movedim(-1, axis)


def e4m3_pos(x):
    """Positive scale rounded up-ish to e4m3 (with a global 2^k pre-scale the kernel would fold into scale-input-d)."""
    k = torch.floor(torch.log2(x.clamp_min(1e-300)))
    m = x / 2.0 ** k                                   # [1, 2)
    m = torch.ceil(m * 8.0) / 8.0
    return m * 2.0 ** k


def pow2_scale(x, target):
    m = float(x.abs().max())
    return 2.0 ** np.floor(np.log2(target / max(m, 1e-30)))


class Scheme:
    def __init__(self, name):
        self.name = name

    def mm(self, a, b, kind):
        """a (rows, K), b (K, N) fp64 (values are fp32-representable) -> fp64."""
        n = self.name
        if n == "fp64":
            return a @ b
        if n == "bf16x3":
            ah, bh = rnd(a, torch.bfloat16), rnd(b, torch.bfloat16)
            al, bl = rnd(a - ah, torch.bfloat16), rnd(b - bh, torch.bfloat16)
            return ah @ bh + al @ bh + ah @ bl
        if n == "bf16":
            return rnd(a, torch.bfloat16) @ rnd(b, torch.bfloat16)
        if n == "f16":
            return rnd(a, torch.float16) @ rnd(b, torch.float16)
        if n == "f16x2a":      # activation split only
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            return ah @ bh + rnd(a - ah, torch.float16) @ bh
        if n.startswith("s15"):
            # scale-input-d variant: both correction products share the scale 2^15 (sa + sb = 15)
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            al, bl = a - ah, b - bh
            sa, sb, sa2, sb2 = (9, 6, 0, 15) if kind == "h" else (10, 5, 0, 15)
            t1 = e4m3(al * 2.0 ** sa) @ e4m3(bh * 2.0 ** sb)
            t2 = e4m3(ah * 2.0 ** sa2) @ e4m3(bl * 2.0 ** sb2)
            return ah @ bh + (t1 + t2) * 2.0 ** -15
        if n == "f16f8a":      # activation-side correction only (K = 64 e4m3 block): 1.5 units
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            al = a - ah
            sb = pow2_scale(b, 256.0)
            sa_lo = 2.0 ** 8 if kind == "h" else 2.0 ** 12
            return ah @ bh + e4m3(al * sa_lo) @ e4m3(bh * sb) / (sa_lo * sb)
        if n == "f16f8w":      # weight-side correction only
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            bl = b - bh
            sbl = pow2_scale(bl, 256.0)
            sa_hi = 2.0 ** -3 if kind == "h" else 2.0 ** 4
            return ah @ bh + e4m3(ah * sa_hi) @ e4m3(bl * sbl) / (sa_hi * sbl)
        if n.startswith("f16f4"):
            # fp16 main product + the two correction products in block-scaled e2m1 (4x rate): 1.5 units
            kind4 = "nvf4" if "nv" in n else "mxf4"
            blk = 16 if kind4 == "nvf4" else 32
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            al, bl = a - ah, b - bh
            t1 = fp4_block(al, 1, blk, kind4) @ fp4_block(bh, 0, blk, kind4)
            t2 = fp4_block(ah, 1, blk, kind4) @ fp4_block(bl, 0, blk, kind4)
            return ah @ bh + t1 + t2
        if n.startswith("f16f8"):
            ah, bh = rnd(a, torch.float16), rnd(b, torch.float16)
            al, bl = a - ah, b - bh
            # static scales: weights per matrix (known at pack time), activations fixed per GEMM kind
            sb = pow2_scale(b, 256.0)
            sbl = pow2_scale(bl, 256.0)
            if kind == "h":      # residual stream operand
                sa_lo, sa_hi = 2.0 ** 8, 2.0 ** -3
            else:                # gated activations in (-1, 1)
                sa_lo, sa_hi = 2.0 ** 12, 2.0 ** 4
            if n == "f16f8_dyn":
                sa_lo, sa_hi = pow2_scale(al, 256.0), pow2_scale(a, 256.0)
            t1 = e4m3(al * sa_lo) @ e4m3(bh * sb) / (sa_lo * sb)
            t2 = e4m3(ah * sa_hi) @ e4m3(bl * sbl) / (sa_hi * sbl)
            return ah @ bh + t1 + t2
        raise ValueError(n)


def wavenet(orc, sch, x, cond):
    n = orc.wn_name
    f64 = torch.float64

    def kern(name):
        return weight_norm_kernel(torch.as_tensor(orc.w[f"{name}/v"], dtype=f64), torch.as_tensor(orc.w[f"{name}/g"], dtype=f64)).to(torch.float32).to(f64)

    def bias(name):
        return torch.as_tensor(orc.w[f"{name}/bias"], dtype=f64)

    h = (x @ kern(f"{n}/start")[0] + bias(f"{n}/start")).to(torch.float32).to(f64)
    out = None
    T = h.shape[0]
    for i, d in enumerate(orc.dilations):
        k = kern(f"{n}/conv1D_{i}")                                  # (3, C, 2C)
        hp = torch.cat([torch.zeros(d, h.shape[1], dtype=f64), h, torch.zeros(d, h.shape[1], dtype=f64)])
        a = torch.cat([hp[0:T], hp[d:d + T], hp[2 * d:2 * d + T]], dim=1)       # (T, 3C)
        z = sch.mm(a, k.reshape(-1, k.shape[2]), "h") + bias(f"{n}/conv1D_{i}") + cond
        zt, zs = torch.split(z, z.shape[1] // 2, dim=1)
        act = (torch.tanh(zt) * torch.sigmoid(zs)).to(torch.float32).to(f64)
        rs = sch.mm(act, kern(f"{n}/res_skip_{i}")[0], "act") + bias(f"{n}/res_skip_{i}")
        if i < orc.n_layers - 1:
            res, skip = torch.split(rs, rs.shape[1] // 2, dim=1)
            h = (h + res).to(torch.float32).to(f64)
            if sch.name.endswith("_hm"):     # residual stream master kept as fp16 hi + e4m3 lo * 2^-9
                hh = rnd(h, torch.float16)
                h = hh + e4m3((h - hh) * 2.0 ** 9) * 2.0 ** -9
        else:
            skip = rs
        out = skip if out is None else out + skip
    return out @ kern(f"{n}/end")[0] + bias(f"{n}/end"), h


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "SPEECH"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    hp = read_config(get_config_file(model))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))
    orc = OracleMBExWN(hp, w, torch.float64)
    mel = torch.as_tensor(synthetic_mel(T, 0)[None], dtype=torch.float64)
    noise = synthetic_noise(T * plan.steps_per_frame, 0)
    f0 = orc.generate_f0(mel)
    pg = orc.pulse_generator(f0.numpy())
    pulse = torch.as_tensor(pg["pulse"]).reshape(1, -1, orc.pulse_channels)
    x = torch.cat([pulse, torch.as_tensor(noise[None], dtype=torch.float64) * orc.sigma], dim=-1)[0]
    _, cond = orc.conditioning(mel)
    cond = cond[0]
    ref, href = wavenet(orc, Scheme("fp64"), x, cond)
    print(f"model {model}: rows {x.shape[0]}, C {orc.C}, |wn_out| peak {float(ref.abs().max()):.3f}, |h| peak {float(href.abs().max()):.2f}")
    for name in ("bf16x3", "f16", "f16f8", "f16f8a", "f16f8w", "s15", "s15_hm", "f16f4nv", "f16f4mx"):
        out, h = wavenet(orc, Scheme(name), x, cond)
        err = (out - ref)
        snr = 10 * np.log10(float((ref ** 2).sum() / (err ** 2).sum()))
        print(f"{name:10s} wn_out max-abs err / peak = {float(err.abs().max() / ref.abs().max()):.2e}   SNR {snr:6.1f} dB   "
              f"h err/peak {float((h - href).abs().max() / href.abs().max()):.2e}")


if __name__ == "__main__":
    main()
