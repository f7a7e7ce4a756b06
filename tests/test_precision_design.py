"""CPU checks of the split-precision (MBEXWN_PREC_F16F8) design: operand plane layout produced by the host packer, and an
emulation of what the tensor cores compute (tools/sim_precision.py) against the fp64 oracle at the north-star tolerance."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from mbexwn_vocoder_b200 import tc_pack  # noqa: E402


def _decode_planes(packed: torch.Tensor, k: int):
    n = packed.shape[0]
    h16 = packed[:, :2 * k].contiguous().view(torch.float16).float()
    f8 = packed[:, 2 * k:].contiguous().view(torch.float8_e4m3fn).float().reshape(n, k // 64, 2, 64)
    return h16, f8[:, :, 0].reshape(n, k), f8[:, :, 1].reshape(n, k)


def test_f16f8_planes_layout_and_reconstruction():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((48, 192)) * 0.05).astype(np.float32)
    hi_shift, lo_shift = 6, 15
    p = tc_pack.f16f8_planes(w, hi_shift, lo_shift)
    assert p.dtype == torch.uint8 and p.shape == (48, 4 * 192)
    h16, hi8, lo8 = _decode_planes(p, 192)
    wt = torch.from_numpy(w)
    assert torch.equal(h16, wt.to(torch.float16).float())
    # first e4m3 block of every 64 columns: the weight itself (scaled), second: its fp16 residual (scaled)
    assert torch.allclose(hi8 * 2.0 ** -hi_shift, wt, rtol=2 ** -4, atol=2.0 ** (-9 - hi_shift))
    resid = wt - h16
    assert torch.allclose(lo8 * 2.0 ** -lo_shift, resid, rtol=2 ** -4, atol=2.0 ** (-9 - lo_shift))
    # fp16 + e4m3 residual carries ~15 mantissa bits
    assert float((h16 + lo8 * 2.0 ** -lo_shift - wt).abs().max()) <= 2.0 ** -16 * float(wt.abs().max())


def test_tc8_shifts_never_saturate_the_weight_plane():
    for wmax in (0.01, 0.3, 3.0, 40.0, 900.0):
        sh = tc_pack.choose_tc8_shifts(wmax, wmax)
        for key in ("tc8_h_lo", "tc8_a_lo"):
            assert wmax * 2.0 ** (tc_pack.CORR_SHIFT - sh[key]) <= tc_pack.E4M3_MAX
            assert sh[key] >= tc_pack.TC8_DEFAULT_SHIFTS[key]


def test_packed_tc8_weights_match_bf16_packing_geometry(speech_setup):
    hp, plan, w = speech_setup
    a = tc_pack.pack_tc_weights(plan, w)
    b, shifts = tc_pack.pack_tc8_weights(plan, w)
    assert shifts == tc_pack.choose_tc8_shifts(0.0, 0.0) or set(shifts) == set(tc_pack.TC8_DEFAULT_SHIFTS)
    for key, t in b.items():
        ref = a[key.replace("/tcf8/", "/tcf/").replace("/tc8/", "/tc/")]
        assert t.shape[0] == ref.shape[0] and t.shape[1] == 2 * ref.shape[1]       # same rows, same bytes per row
        k = ref.shape[1] // 2
        h16, _, _ = _decode_planes(t, k)
        assert torch.allclose(h16, ref[:, :k].float(), rtol=2 ** -7, atol=1e-6)     # fp16 vs bf16 rounding of the same matrix


def test_emulated_f16f8_wavenet_meets_the_fp32_tolerance():
    """What the kernel computes (fp16 product + 2^-15 (e4m3 x e4m3 + e4m3 x e4m3), residual stream kept as fp16 + e4m3)
    against the fp64 oracle: north-star bar for the fp32-accurate path is 1e-4 of peak and 60 dB."""
    import sim_precision as sp
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise
    hp = read_config(get_config_file("SPEECH"))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp["synthetic_weights"]["seed"]))
    orc = OracleMBExWN(hp, w, torch.float64)
    T = 12
    mel = torch.as_tensor(synthetic_mel(T, 0)[None], dtype=torch.float64)
    f0 = orc.generate_f0(mel)
    pulse = torch.as_tensor(orc.pulse_generator(f0.numpy())["pulse"]).reshape(1, -1, orc.pulse_channels)
    noise = torch.as_tensor(synthetic_noise(T * plan.steps_per_frame, 0)[None], dtype=torch.float64) * orc.sigma
    x = torch.cat([pulse, noise], dim=-1)[0]
    cond = orc.conditioning(mel)[1][0]
    ref, _ = sp.wavenet(orc, sp.Scheme("fp64"), x, cond)
    out, _ = sp.wavenet(orc, sp.Scheme("s15_hm"), x, cond)
    err = out - ref
    assert float(err.abs().max() / ref.abs().max()) <= 1e-4
    assert 10 * np.log10(float((ref ** 2).sum() / (err ** 2).sum())) >= 60.0
    # plain bf16 (the separately stated bf16 path) is ~50 dB worse
    out16, _ = sp.wavenet(orc, sp.Scheme("bf16"), x, cond)
    e16 = out16 - ref
    assert 10 * np.log10(float((ref ** 2).sum() / (e16 ** 2).sum())) >= 35.0


def test_chunked_gate_permutation_of_the_fused_layer_kernel():
    """W1 rows for csrc/k_wavenet_layer.cu: per 16-channel chunk [16 tanh rows | their 16 sigmoid partners] -- a permutation of the
    reference's [tanh(C) | sigmoid(C)] columns (custom_AE_layers.py:309-321) with the padding rows masked."""
    for c, cpad in ((320, 320), (340, 384)):
        col, ok = tc_pack.gate_permutation_chunked(c, cpad)
        assert col.shape == (2 * cpad,) and sorted(col[ok]) == list(range(2 * c))
        n = np.arange(2 * cpad)
        chunk, within = n // 32, n % 32
        assert np.array_equal(col[ok] % c, (16 * chunk + within % 16)[ok])          # tanh row and sigmoid row: same channel
        assert np.array_equal(col[ok] >= c, (within >= 16)[ok])                      # sigmoid half = second 16 rows of a chunk
