"""GPU tests of the tcgen05 tap-GEMM kernel and the tensor-core WaveNet path."""
import numpy as np
import pytest
import torch

from oracle.forward import OracleMBExWN, synthetic_mel, synthetic_noise

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(speech_setup):
    from mbexwn_vocoder_b200.engine import Engine
    hp, plan, w = speech_setup
    eng = Engine(plan, w, device=0)
    yield eng
    eng.close()


def _ref_gemm(a, b, kblocks):
    rows = a.shape[0]
    af, bf = a.float().cpu().numpy().astype(np.float64), b.float().cpu().numpy().astype(np.float64)
    out = np.zeros((rows, b.shape[0]))
    for a_col, shift, b_col in kblocks:
        blk = np.zeros((rows, 64))
        lo, hi = max(0, -shift), min(rows, rows - shift)
        blk[lo:hi] = af[lo + shift:hi + shift, a_col:a_col + 64]
        out += blk @ bf[:, b_col:b_col + 64].T
    return out


@pytest.fixture(params=[1, 2], ids=["cta_group1", "cta_group2"])
def cg(request, engine):
    engine.set_option("tc_cta_group", request.param)
    yield request.param
    engine.set_option("tc_cta_group", 1)


@pytest.mark.parametrize("rows,n,kblocks", [
    (128, 128, [(0, 0, 0)]),
    (256, 128, [(0, 0, 0), (64, 0, 64)]),
    (300, 256, [(0, 0, 0), (64, 0, 64), (0, -3, 128), (64, 5, 192)]),
    (200, 32, [(0, 0, 0), (64, 1, 64)]),
    (513, 240, [(0, -1, 0), (0, 0, 64), (0, 1, 128)]),
    (1000, 320, [(64 * i, s, 64 * (3 * i + j)) for i in range(3) for j, s in enumerate((-8, 0, 8))]),
])
def test_tap_gemm_exact(engine, cg, rows, n, kblocks):
    g = torch.Generator(device="cpu").manual_seed(rows + n)
    a_cols = max(k[0] for k in kblocks) + 64
    b_cols = max(k[2] for k in kblocks) + 64
    # small integers are exact in bf16 and in the fp32 accumulator: the result must match bit for bit
    a = torch.randint(-4, 5, (rows, a_cols), generator=g).to(torch.bfloat16).cuda()
    b = torch.randint(-4, 5, (n, b_cols), generator=g).to(torch.bfloat16).cuda()
    out = engine.tc_gemm(a, b, np.array(kblocks)).cpu().numpy()
    ref = _ref_gemm(a, b, kblocks)
    assert np.array_equal(out, ref.astype(np.float32))


def test_tap_gemm_random(engine, cg):
    g = torch.Generator(device="cpu").manual_seed(7)
    rows, n = 777, 640
    kblocks = [(64 * c, (t - 1) * 4, t * 320 + 64 * c) for t in range(3) for c in range(5)]
    a = torch.randn(rows, 320, generator=g).to(torch.bfloat16).cuda()
    b = (torch.randn(n, 960, generator=g) * 0.05).to(torch.bfloat16).cuda()
    out = engine.tc_gemm(a, b, np.array(kblocks)).cpu().numpy()
    ref = _ref_gemm(a, b, kblocks)
    assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max()


def _pack_f16f8(x16, first8, second8):
    """rows of [fp16 | per 64 columns: e4m3 first8 (64), e4m3 second8 (64)] as a uint8 matrix (the MBEXWN_PREC_F16F8 operand
    layout: activations store (lo8, hi8), weights (hi8, lo8))."""
    n, k = x16.shape
    a = first8.to(torch.float8_e4m3fn).view(torch.uint8).reshape(n, k // 64, 64)
    b = second8.to(torch.float8_e4m3fn).view(torch.uint8).reshape(n, k // 64, 64)
    return torch.cat((x16.to(torch.float16).view(torch.uint8), torch.stack((a, b), dim=2).reshape(n, 2 * k)), dim=1).contiguous()


def _ref_gemm_planes(a, b, kblocks):
    rows = a.shape[0]
    out = np.zeros((rows, b.shape[0]))
    a, b = a.double().numpy(), b.double().numpy()
    for a_col, shift, b_col in kblocks:
        blk = np.zeros((rows, 64))
        lo, hi = max(0, -shift), min(rows, rows - shift)
        blk[lo:hi] = a[lo + shift:hi + shift, a_col:a_col + 64]
        out += blk @ b[:, b_col:b_col + 64].T
    return out


@pytest.mark.parametrize("rows,n,kblocks", [
    (128, 128, [(0, 0, 0)]),
    (300, 256, [(0, 0, 0), (64, 0, 64), (0, -3, 128), (64, 5, 192)]),
    (200, 32, [(0, 0, 0), (64, 1, 64)]),
    (513, 240, [(0, -1, 0), (0, 0, 64), (0, 1, 128)]),
    (1000, 320, [(64 * i, s, 64 * (3 * i + j)) for i in range(3) for j, s in enumerate((-8, 0, 8))]),
])
def test_tap_gemm_f16f8_exact(engine, cg, rows, n, kblocks):
    """fp16 main product + 2^-15 (lo8 x hi8 + hi8 x lo8) with small integers: every partial sum is exact in fp32."""
    g = torch.Generator(device="cpu").manual_seed(rows * 3 + n)
    a_cpad = max(k[0] for k in kblocks) + 64
    b_k = max(k[2] for k in kblocks) + 64
    amp = 4 if len(kblocks) <= 4 else 2          # keep every partial sum below 2^9 (15 fractional bits + 9 = fp32's 24)
    ri = lambda *shape: torch.randint(-amp, amp + 1, shape, generator=g).float()
    a16, alo, ahi = ri(rows, a_cpad), ri(rows, a_cpad), ri(rows, a_cpad)
    b16, bhi, blo = ri(n, b_k), ri(n, b_k), ri(n, b_k)
    out = engine.tc_gemm_f16f8(_pack_f16f8(a16, alo, ahi).cuda(), _pack_f16f8(b16, bhi, blo).cuda(), np.array(kblocks)).cpu().numpy()
    ref = _ref_gemm_planes(a16, b16, kblocks) + 2.0 ** -15 * (_ref_gemm_planes(alo, bhi, kblocks) + _ref_gemm_planes(ahi, blo, kblocks))
    assert np.array_equal(out, ref.astype(np.float32))


def _snr_db(ref, test):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum(err ** 2), 1e-300))


@pytest.fixture(params=[0, 1], ids=["two_launch", "fused_layer"])
def fused(request, engine):
    """tc_fused = 1 forces the one-kernel-per-layer path (k_wavenet_layer.cu) also for batches this short (CTA pairs only)."""
    engine.set_option("tc_fused", request.param)
    yield request.param
    engine.set_option("tc_fused", 1)


@pytest.mark.parametrize("precision,tol,snr", [("bf16x3", 1e-4, 60.0), ("f16f8", 1e-4, 60.0), ("bf16", 5e-2, 35.0)])
def test_wavenet_tc_parity(engine, cg, fused, speech_setup, precision, tol, snr):
    hp, plan, w = speech_setup
    oracle = OracleMBExWN(hp, w, torch.float32)
    lengths = [23, 57, 10]
    mels = [synthetic_mel(t, i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, i) for i, t in enumerate(lengths)]
    f0 = [oracle.generate_f0(torch.as_tensor(m[None])).numpy()[0] for m in mels]
    out, tp = engine.forward(mels, noise=noise, f0=f0, precision=precision, taps=["index", "wn_out", "subbands", "excitation"])
    for u, t in enumerate(lengths):
        ref = oracle.forward(mels[u][None], noise[u][None], f0_override=f0[u][None])
        assert np.array_equal(tp["index"][u], ref["index"][0])
        for st in ["wn_out", "subbands", "excitation"]:
            r = np.asarray(ref[st][0]).reshape(-1)
            e = np.abs(tp[st][u] - r).max() / np.abs(r).max()
            print(f"{precision} utt {u} {st}: max|err|/peak {e:.3e}")
            assert e <= tol, st
        s = _snr_db(ref["waveform"][0], out[u])
        e = np.abs(out[u] - ref["waveform"][0]).max() / np.abs(ref["waveform"][0]).max()
        print(f"{precision} utt {u} waveform: max|err|/peak {e:.3e}  SNR {s:.1f} dB")
        assert s >= snr and e <= tol


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3", "bf16"])
def test_fused_layer_kernel_equals_two_launch_path(engine, speech_setup, precision):
    """One persistent kernel per layer (gate tiles of M tile m, res tiles of M tile m - 1, activations through the L2 scratch,
    residual stream ping-pong) against the gate + res/skip launches on a batch that gives every CTA pair several M tiles
    and a ragged tail: the same products, summed in a different K order -- fp32 rounding apart (5e-5 of peak), every buffer agrees."""
    hp, plan, w = speech_setup
    lengths = [400, 380, 17, 400, 211, 400, 1, 400, 399, 400, 2, 400, 400, 333]
    mels = [synthetic_mel(t, 40 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 40 + i) for i, t in enumerate(lengths)]
    engine.set_option("tc_cta_group", 2)
    res = []
    for mode in (0, 1):
        engine.set_option("tc_fused", mode)
        out, tp = engine.forward(mels, noise=noise, precision=precision, taps=["wn_out", "excitation"])
        res.append((out, tp))
        res[-1] += (int(engine.lib.mbexwn_last_launch_count(engine._handle)),)
    # one launch per layer against two: the fused kernel must really have run
    assert res[0][2] - res[1][2] == plan.wavenet.n_layers, (res[0][2], res[1][2])
    engine.set_option("tc_fused", 1)
    engine.set_option("tc_cta_group", 1)
    for u in range(len(lengths)):
        assert np.all(np.isfinite(res[1][0][u]))
        tol = 3e-2 if precision == "bf16" else 5e-5          # bf16 re-rounds the activations of every layer
        for a, b, what in ((res[0][1]["wn_out"][u], res[1][1]["wn_out"][u], "wn_out"), (res[0][0][u], res[1][0][u], "waveform")):
            assert np.abs(a - b).max() <= tol * max(np.abs(a).max(), 1e-30), f"{what} of utterance {u}"


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_fused_layer_slab_views_equal_per_tap_loads(engine, speech_setup, precision):
    """The dilated taps of the fused kernel read ONE slab per 64-channel block through row-shifted MMA descriptor views
    (SWIZZLE_128B with the descriptor's base offset); with "tc_slab" = 0 every tap gets its own TMA tile.  Same products in
    the same order: the outputs must be bit-identical -- for every dilation of the stack (shifts of 1, 2, 4 and 8 rows)."""
    hp, plan, w = speech_setup
    lengths = [400, 150, 1, 400, 37, 400, 400, 260]
    mels = [synthetic_mel(t, 60 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 60 + i) for i, t in enumerate(lengths)]
    engine.set_option("tc_cta_group", 2)
    engine.set_option("tc_fused", 1)
    res = []
    for slab in (0, 1):
        engine.set_option("tc_slab", slab)
        out, tp = engine.forward(mels, noise=noise, precision=precision, taps=["wn_out"])
        res.append((out, tp))
    engine.set_option("tc_fused", 1)
    engine.set_option("tc_cta_group", 1)
    for u in range(len(lengths)):
        assert np.all(np.isfinite(res[1][0][u]))
        assert np.array_equal(res[0][1]["wn_out"][u], res[1][1]["wn_out"][u]), f"wn_out of utterance {u}"
        assert np.array_equal(res[0][0][u], res[1][0][u]), f"waveform of utterance {u}"


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_fused_layer_clusters_of_four_equal_pairs(engine, speech_setup, precision):
    """"tc_cluster" = 4: two CTA pairs of a cluster take turns loading the weight tiles and multicast them to each other
    (the second pair of the last cluster may run one step past the last row).  Same products in the same order as with one
    pair per cluster: bit-identical outputs, for a batch large enough that every pair gets several 256-row tiles."""
    hp, plan, w = speech_setup
    lengths = [400] * 14 + [150, 1, 37, 260, 333]
    mels = [synthetic_mel(t, 80 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 80 + i) for i, t in enumerate(lengths)]
    engine.set_option("tc_cta_group", 2)
    engine.set_option("tc_fused", 1)
    res, used = [], []
    for cluster in (2, 4):
        engine.set_option("tc_cluster", cluster)
        out, tp = engine.forward(mels, noise=noise, precision=precision, taps=["wn_out"])
        res.append((out, tp))
        used.append(engine.get_info("tc_last_cluster"))
    quads = engine.get_info("tc_max_quads")
    engine.set_option("tc_cluster", 2)
    engine.set_option("tc_cta_group", 1)
    print(f"clusters of 4 resident at once: {quads}; cluster sizes used: {used}")
    assert used[0] == 2
    if quads * 4 >= 128:
        assert used[1] == 4, "the batch is large enough for clusters of 4"
    for u in range(len(lengths)):
        assert np.all(np.isfinite(res[1][0][u]))
        assert np.array_equal(res[0][1]["wn_out"][u], res[1][1]["wn_out"][u]), f"wn_out of utterance {u}"
        assert np.array_equal(res[0][0][u], res[1][0][u]), f"waveform of utterance {u}"


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_fused_layer_interleaved_tile_order_is_bit_identical(engine, speech_setup, precision):
    """"tc_interleave" = 1 (default) runs the res tiles of M tile j - 1 behind the first gate tile of M tile j (cut at a
    64-channel block: 256 + 192 + 192 columns instead of 224 + 224 + 192) so that they find their operands in the L2, and
    "tc_discard" drops the scratch rows from the L2 once they have been read.  Every output element sums the same products in
    the same order: all four combinations agree bit for bit."""
    hp, plan, w = speech_setup
    lengths = [400, 150, 1, 400, 37, 400, 400, 260]
    mels = [synthetic_mel(t, 90 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 90 + i) for i, t in enumerate(lengths)]
    engine.set_option("tc_cta_group", 2)
    engine.set_option("tc_fused", 1)
    res = []
    for il, discard in ((0, 0), (1, 1), (0, 1), (1, 0)):
        engine.set_option("tc_interleave", il)
        engine.set_option("tc_discard", discard)                   # dead scratch rows dropped from the L2: must not change a bit
        res.append(engine.forward(mels, noise=noise, precision=precision, taps=["wn_out"]))
    engine.set_option("tc_interleave", 1)
    engine.set_option("tc_discard", 1)
    engine.set_option("tc_cta_group", 1)
    for u in range(len(lengths)):
        assert np.all(np.isfinite(res[1][0][u]))
        for k in (1, 2, 3):
            assert np.array_equal(res[0][1]["wn_out"][u], res[k][1]["wn_out"][u]), f"wn_out of utterance {u}, variant {k}"
            assert np.array_equal(res[0][0][u], res[k][0][u]), f"waveform of utterance {u}, variant {k}"


@pytest.mark.parametrize("lengths", [[40], [23, 57, 10, 1, 2]])
def test_subnets_tc_parity(engine, cg, speech_setup, lengths):
    """F0 net, VTF net and conditioning conv as 3-product bf16 tap-GEMMs (mirrored pad rows, sub-pixel unfold, PReLU
    epilogue) against the fp32 oracle: every stage within 1e-4 of its peak, also for 1- and 2-frame utterances."""
    hp, plan, w = speech_setup
    oracle = OracleMBExWN(hp, w, torch.float32)
    mels = [synthetic_mel(t, 20 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 20 + i) for i, t in enumerate(lengths)]
    out, tp = engine.forward(mels, noise=noise, precision="bf16x3", taps=["F0", "cond", "ceps"])
    assert engine.lib.mbexwn_last_launch_count(engine._handle) > 0
    for u, t in enumerate(lengths):
        ref = oracle.forward(mels[u][None], noise[u][None])
        ref["cond"] = ref["cond_lo"]
        for st in ("F0", "cond", "ceps"):
            r = np.asarray(ref[st][0]).reshape(-1)
            e = np.abs(tp[st][u].reshape(-1) - r).max() / np.abs(r).max()
            print(f"subnets tc utt {u} T={t} {st}: max|err|/peak {e:.3e}")
            assert e <= 1e-4, st


def test_wavenet_tc_parity_c340():
    """MW-VO-FD geometry: C = 340 is padded to 384 channels inside the tensor-core path (zero weights)."""
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.engine import Engine
    from mbexwn_vocoder_b200.plan import build_plan
    hp = read_config(get_config_file("VOICE"))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=13)
    eng = Engine(plan, w, device=0)
    oracle = OracleMBExWN(hp, w, torch.float32)
    lengths = [31, 12]
    mels = [synthetic_mel(t, 10 + i) for i, t in enumerate(lengths)]
    noise = [synthetic_noise(t * plan.steps_per_frame, 10 + i) for i, t in enumerate(lengths)]
    # wide synthetic F0 sweep (45..1400 Hz) with vibrato to exercise every wavetable of the grid
    f0 = []
    for i, t in enumerate(lengths):
        n = t * plan.pulse_per_frame
        x = np.linspace(0, 1, n)
        f0.append((45.0 * (1400.0 / 45.0) ** x * (1 + 0.03 * np.sin(2 * np.pi * 5.5 * np.arange(n) / 8000.0))).astype(np.float32))
    eng.set_option("tc_cta_group", 2)
    for precision, tol, snr, fused in (("bf16x3", 1e-4, 60.0, 0), ("f16f8", 1e-4, 60.0, 0), ("f16f8", 1e-4, 60.0, 1),
                                       ("bf16x3", 1e-4, 60.0, 1), ("fp32", 1e-4, 60.0, 0)):
        eng.set_option("tc_fused", fused)
        out, tp = eng.forward(mels, noise=noise, f0=f0, precision=precision, taps=["index", "phase", "pulse", "wn_out"])
        for u in range(len(lengths)):
            ref = oracle.forward(mels[u][None], noise[u][None], f0_override=f0[u][None])
            assert np.array_equal(tp["index"][u], ref["index"][0])
            assert np.array_equal(tp["phase"][u], ref["phase"][0])
            for st in ("pulse", "wn_out"):
                r = np.asarray(ref[st][0]).reshape(-1)
                e = np.abs(tp[st][u] - r).max() / np.abs(r).max()
                print(f"C340 {precision} utt {u} {st}: {e:.3e}")
                assert e <= tol
            s = _snr_db(ref["waveform"][0], out[u])
            print(f"C340 {precision} utt {u} waveform SNR {s:.1f} dB")
            assert s >= snr
    eng.close()
