"""Host-side logic that runs without a GPU: config reader, model registry, plan, batch geometry, C-ABI surface."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import yaml

import mbexwn_vocoder_b200 as pkg
from mbexwn_vocoder_b200 import _cabi, config as cutils, sched
from mbexwn_vocoder_b200 import weights as W
from mbexwn_vocoder_b200.plan import ACT_PRELU, PAD_SYMMETRIC, PAD_ZERO, build_plan, subnet_program

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- registry / config (MBExWN_NVoc/__init__.py, config_utils.py) ---------------------------------------------
def test_list_models_matches_reference_registry():
    m = pkg.list_models()
    assert set(m) == {"SING", "SPEECH", "VOICE"}
    assert "WNCHA340" in m["VOICE"][0] and "WNCHA320" in m["SPEECH"][0]
    m["SING"].clear()
    assert pkg.list_models()["SING"], "list_models must return a copy"
    assert pkg.mbexwn_version == (1, 2, 3)


def test_get_config_file_lookup(tmp_path):
    for key in ("SPEECH", "SING", "VOICE", "MW-SP-FD", "SPEECH/MBExWN_SIIConv_V71g_SPEECH"):
        assert os.path.exists(pkg.get_config_file(key))
    with pytest.raises(FileNotFoundError):
        pkg.get_config_file("NO_SUCH_MODEL")
    with pytest.raises(FileNotFoundError):
        pkg.get_config_file(str(tmp_path))                    # directory without config.yaml
    (tmp_path / "config.yaml").write_text("a: 1\n")
    assert pkg.get_config_file(str(tmp_path)) == os.path.join(str(tmp_path), "config.yaml")


def test_read_config_defaults_includes_and_types(tmp_path, monkeypatch):
    (tmp_path / "inc.yaml").write_text("sub: {x: 7}\n")
    (tmp_path / "c.yaml").write_text(
        "top:\n  __defaults__: {a: 1, b: 2}\n  b: 5\n"
        "lst:\n  - __defaults__: {k: 3}\n  - {name: p}\n  - {name: q, k: 9}\n"
        "ftype: np.float32\nnone_val: None\n"
        "inc: <@CONFIG_DIR@/inc.yaml:sub:x>\n"
        "home: ~/x\nenv: $MBX_TEST_VAR/y\n")
    monkeypatch.setenv("MBX_TEST_VAR", "/zz")
    c = cutils.read_config(str(tmp_path / "c.yaml"), config_base_dir=str(tmp_path))
    assert c["top"] == {"a": 1, "b": 5}
    assert c["lst"] == [{"name": "p", "k": 3}, {"name": "q", "k": 9}]
    assert c["ftype"] is np.float32 and c["none_val"] is None
    assert c["inc"] == 7 and c["env"] == "/zz/y" and c["home"] == os.path.expanduser("~/x")


def test_read_config_rejects_double_defaults(tmp_path):
    (tmp_path / "c.yaml").write_text("lst:\n  - __defaults__: {k: 3}\n  - __defaults__: {k: 4}\n  - {n: 1}\n")
    with pytest.raises(RuntimeError):
        cutils.read_config(str(tmp_path / "c.yaml"))


# ---- plan (MBExWN.__init__ rate algebra and sub-net grammar) ----------------------------------------------------
def test_plan_rates_and_wavenet_geometry(speech_setup):
    hp, plan, w = speech_setup
    assert (plan.pulse_per_frame, plan.steps_per_frame, plan.pulse_rate) == (100, 20, 8000.0)
    assert plan.wavenet.dilations == [1, 2, 4, 8, 1, 2, 4, 8] and plan.wavenet.cond_conv_up == 2
    assert plan.wavenet.c_in == 6 and plan.stft_win == 1200 and plan.fft_size == 2048
    assert plan.wavetables.tables.shape[0] == plan.wavetables.n_period + 1
    assert abs(plan.filter_max_log_range - 40 / 8.685889638) < 1e-6


def test_subnet_grammar_matches_reference_layer_list():
    ops, ups = subnet_program([[3, 128], [3, 64, 2], [3, 32, "L5"], ["L", 2]], "X", 80, 1, 1, 3, 0.02, None, True)
    kinds = [(o.kind, o.up, o.act) for o in ops]
    # conv+PReLU | sub-pixel conv+PReLU | conv, LinInterp+PReLU | bare LinInterp (no act, Q2) | final 1x1
    assert kinds == [("conv", 1, ACT_PRELU), ("conv", 1, ACT_PRELU), ("conv", 1, 0), ("lininterp", 5, ACT_PRELU),
                     ("lininterp", 2, 0), ("conv", 1, 3)]
    assert ups == 10                                         # quirk Q2: the bare "L" entry does not count
    assert ops[0].conv.pad_mode == PAD_SYMMETRIC and (ops[0].conv.pad_l, ops[0].conv.pad_r) == (1, 1)
    assert ops[1].conv.pad_mode == PAD_ZERO and ops[1].conv.subpixel == 2 and ops[1].conv.cout == 128
    assert ops[-1].rate_in == 20
    with pytest.raises(RuntimeError):
        subnet_program([[3, 8, 3]], "X", 80, 1, 1, 3, 0.02, 100, True)      # 100 not reachable from 3


def test_plan_config_errors(speech_setup):
    import copy
    hp, _, _ = speech_setup
    bad = copy.deepcopy(hp)
    bad["mbexwn_config"]["pulse_channels"] = 4              # 8000/4*15 != 24000 (custom_pulsed_generator.py:344)
    with pytest.raises(RuntimeError):
        build_plan(bad, finalize=False)
    bad = copy.deepcopy(hp)
    del bad["mbexwn_config"]
    with pytest.raises(NotImplementedError):
        build_plan(bad)                                      # models.py:31
    bad = copy.deepcopy(hp)
    bad["use_tf25_compatible_implementation"] = False
    with pytest.raises(NotImplementedError):
        build_plan(bad)                                      # custom_pulsed_generator.py:272
    bad = copy.deepcopy(hp)
    bad["mbexwn_config"]["pp_mod_subnet"]["activation"] = "relu"
    with pytest.raises(RuntimeError):
        build_plan(bad, finalize=False)                      # custom_AE_layers.py:156


def test_synthetic_weights_shapes_and_fold(speech_setup):
    hp, plan, w = speech_setup
    W.check(plan, w)
    name = plan.conv_layers()[0].name
    wf, b = W.folded(w, name)
    v, g = w[f"{name}/v"].astype(np.float64), w[f"{name}/g"].astype(np.float64)
    assert np.allclose(np.sqrt((wf.astype(np.float64) ** 2).sum(axis=(0, 1))), g, rtol=1e-6)
    plain = W.init_synthetic(plan, seed=1, lively=False)
    assert np.all(plain[f"{name}/bias"] == 0)
    assert np.allclose(plain[f"{name}/g"], np.linalg.norm(plain[f"{name}/v"].reshape(-1, plain[f"{name}/v"].shape[2]), axis=0))
    cond = [l for l in plan.conv_layers() if l.name.endswith("/cond_")][0]
    vc = plain[f"{cond.name}/v"].reshape(cond.k, cond.cin, cond.cb_free, -1)
    assert np.array_equal(vc[:, :, 0], vc[:, :, 1])          # checkerboard-free init (conv_layers.py:73-77)


# ---- batch geometry ------------------------------------------------------------------------------------------
def test_frame_grid_layout():
    L = sched.make_layout([3, 12, 1], halo=2, pulse_per_frame=100)
    assert L.n_frames == 2 + 3 + 2 + 12 + 2 + 1 + 2
    assert L.utt_begin.tolist() == [2, 7, 21] and L.utt_end.tolist() == [5, 19, 22]
    assert L.frame_utt.tolist().count(-1) == 8 and L.frame_utt[7] == 1 and L.frame_utt[6] == -1
    assert L.chunk_first.tolist() == [0, 1, 3, 4] and L.n_chunks == 4
    grid = L.scatter([np.full((t * 2, 1), i + 1.0) for i, t in enumerate([3, 12, 1])], 2, np.zeros((L.n_frames * 2, 1)))
    parts = L.gather(grid, 2)
    assert [p.shape[0] for p in parts] == [6, 24, 2] and all(np.all(p == i + 1) for i, p in enumerate(parts))
    assert grid.sum() == 6 * 1 + 24 * 2 + 2 * 3
    with pytest.raises(ValueError):
        sched.make_layout([3, 0], 1, 100)
    with pytest.raises(ValueError):
        sched.make_layout([], 1, 100)


def test_lpt_sharding_is_balanced_and_deterministic():
    rng = np.random.default_rng(1)
    lengths = rng.integers(80, 2401, size=8192)
    for n in (2, 4, 8):
        shards = sched.lpt_shards(lengths, n)
        assert sorted(i for s in shards for i in s) == list(range(8192))
        loads = [int(lengths[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= 2400
        assert shards == sched.lpt_shards(lengths, n)


# ---- C-ABI surface -----------------------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mbexwn.h")).read()
    return sorted(set(re.findall(r"MBEXWN_API\s+[\w\s\*]+?\b(mbexwn_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_cabi.SYMBOLS)


def test_shared_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.lib_path()):
        subprocess.run(["bash", os.path.join(ROOT, "mbexwn_vocoder_b200", "csrc", "build.sh")], check=True)
    lib = _cabi.load()
    for name in _declared_symbols():
        assert getattr(lib, name) is not None
    assert lib.mbexwn_abi_version() == _cabi.ABI_VERSION


def test_struct_layout_matches_header_sizes():
    """sizeof() of the ctypes mirrors must equal what the C compiler sees for include/mbexwn.h."""
    src = ('#include <stdio.h>\n#include "mbexwn.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(mbexwn_op_t), '
           'sizeof(mbexwn_config_t), sizeof(mbexwn_batch_t), sizeof(mbexwn_analysis_config_t), '
           'sizeof(mbexwn_analysis_batch_t));return 0;}\n')
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(_cabi.Op), ctypes.sizeof(_cabi.Config), ctypes.sizeof(_cabi.Batch),
                                     ctypes.sizeof(_cabi.AnalysisConfig), ctypes.sizeof(_cabi.AnalysisBatch)]


def test_no_cpu_fallback_without_gpu(speech_setup):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mbexwn_vocoder_b200.engine import Engine
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    hp, plan, w = speech_setup
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(plan, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MELInverter("SPEECH")
    lib = _cabi.load()
    h = ctypes.c_void_p()
    from mbexwn_vocoder_b200.engine import make_config
    assert lib.mbexwn_create(ctypes.byref(make_config(plan)), ctypes.byref(h)) == _cabi.ERR_CUDA


def test_missing_weights_fail_hard_unless_asked_for_synthetic_ones(monkeypatch, speech_setup):
    """The reference fails in load_weights when a model directory has no checkpoint (mel_inverter.py:203-210); so does
    resolve_weights -- random weights need an explicit opt-in (argument or MBEXWN_SYNTHETIC_WEIGHTS=1) and announce themselves."""
    from mbexwn_vocoder_b200 import get_config_file
    from mbexwn_vocoder_b200.mel_inverter import resolve_weights
    hp, plan, w = speech_setup
    model_dir = os.path.dirname(get_config_file("SPEECH"))
    monkeypatch.delenv("MBEXWN_SYNTHETIC_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError, match="allow_synthetic_weights"):
        resolve_weights(model_dir, plan, hp)
    got = resolve_weights(model_dir, plan, hp, allow_synthetic_weights=True)
    assert set(got) == set(w)
    monkeypatch.setenv("MBEXWN_SYNTHETIC_WEIGHTS", "1")
    assert set(resolve_weights(model_dir, plan, hp)) == set(w)
    with pytest.raises(FileNotFoundError):
        resolve_weights(model_dir, plan, hp, allow_synthetic_weights=False)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mbexwn_vocoder_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_scale_mel_host_logic():
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    inv = MELInverter(None)
    inv.hop_size, inv._srate, inv.fft_size, inv.fmin, inv.fmax, inv.mel_channels = 300, 24000, 2048, 0, 12000, 80
    rng = np.random.default_rng(0)
    mell = np.log(rng.uniform(1e-4, 1.0, size=(80, 25))).astype(np.float32)
    d = {"mell": mell, "sr": 24000, "hoplen": 300, "nfft": 2048, "fmin": 0, "fmax": 12000}
    out = inv.scale_mel(d)
    assert out.shape == (1, 25, 80) and out.dtype == np.float32
    assert np.allclose(out[0], np.log(np.exp(mell.T) + 1e-5), atol=1e-5)
    with pytest.raises(RuntimeError):
        inv.scale_mel(dict(d, fmin=50))
    with pytest.raises(RuntimeError):
        inv.scale_mel(dict(d, fmax=8000))
    half = inv.scale_mel(dict(d, nfft=1024))                  # fft-size rescale (mel_inverter.py:85-87)
    assert np.allclose(half[0], np.log(2 * np.exp(mell.T) + 1e-5), atol=1e-5)
    slow = inv.scale_mel(dict(d, hoplen=600))                 # hop re-interpolation (mel_inverter.py:117-146)
    assert slow.shape[1] == 49


# ---- multi-process host logic (gloo, world size 2) ----------------------------------------------------------
_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mbexwn_vocoder_b200 import sched
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lengths = np.random.default_rng(7).integers(80, 2401, size=64)
mine = sched.lpt_shards(lengths, 2)[rank]
layout = sched.make_layout(lengths[mine], 1, 100)
# stand-in for the per-rank forward: every utterance yields T*300 samples; final step = host gather of sizes
sizes = torch.tensor([int(lengths[i]) * 300 for i in mine] + [0] * (64 - len(mine)), dtype=torch.int64)
idx = torch.tensor(list(mine) + [-1] * (64 - len(mine)), dtype=torch.int64)
gs, gi = [torch.zeros_like(sizes) for _ in range(2)], [torch.zeros_like(idx) for _ in range(2)]
dist.all_gather(gs, sizes); dist.all_gather(gi, idx)
t = torch.tensor([float(layout.n_frames)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    got = {int(i): int(s) for g_i, g_s in zip(gi, gs) for i, s in zip(g_i, g_s) if i >= 0}
    assert got == {i: int(lengths[i]) * 300 for i in range(64)}, "gathered outputs do not cover the batch"
    print("OK", int(t.item()))
dist.destroy_process_group()
'''


def test_two_rank_sharding_and_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert outs[0][0].startswith("OK")


# ---- chunked long-form synthesis: host logic (long_form.py) -----------------------------------------------------
def test_long_form_windows_tile_the_signal():
    from mbexwn_vocoder_b200.long_form import plan_windows
    for T, chunk, ctx, align in [(173, 40, 17, 10), (48000, 400, 23, 10), (5, 400, 23, 10), (400, 400, 23, 10)]:
        wins = plan_windows(T, chunk, ctx, align)
        assert wins[0].core0 == 0 and wins[-1].core1 == T
        for a, b in zip(wins, wins[1:]):
            assert a.core1 == b.core0
        for w in wins:
            assert w.start % align == 0 and 0 <= w.start <= w.core0 < w.core1 <= w.stop <= T
            assert w.core0 - w.start >= min(ctx, w.core0) and w.stop - w.core1 >= min(ctx, T - w.core1)


def test_phase_carry_restates_the_reference_offsets(speech_setup):
    """phase_run_before_chunks continues stable_cumsum_and_wrap (tf_wavetable.py:429-492) across a cut on a chunk boundary:
    the oracle phase of a whole utterance equals the phase of its tail computed with the carried running sum."""
    import torch
    from mbexwn_vocoder_b200.long_form import phase_run_before_chunks
    from oracle.forward import OracleMBExWN
    hp, plan, w = speech_setup
    orc = OracleMBExWN(hp, w, torch.float32)
    n, cut = 7300, 3000
    rng = np.random.default_rng(3)
    f0 = (80.0 + 400.0 * rng.random(n)).astype(np.float32)
    v = (f0 / np.float32(plan.pulse_rate))[None]
    whole = orc.stable_cumsum_and_wrap(v)[0]
    run = phase_run_before_chunks(f0, plan.pulse_rate, 1000)
    assert run.shape == (8,) and run[0] == 0
    # tail alone: in-chunk cumsum + (carry + running sum of its own wrapped totals), wrapped like the reference
    tail = np.zeros(5000, dtype=np.float32)
    tail[:n - cut] = v[0, cut:]
    cum = np.cumsum(tail.reshape(5, 1000), axis=1, dtype=np.float32)
    tot = np.mod(cum[:, -1], np.float32(1))
    offs = np.empty(5, dtype=np.float32)
    acc = run[cut // 1000]
    for j in range(5):
        offs[j] = acc
        acc = np.float32(acc + tot[j])
    phase = np.mod(cum + np.mod(offs, np.float32(1))[:, None], np.float32(1)).reshape(-1)[:n - cut]
    assert np.array_equal(phase, whole[cut:])


def test_long_form_context_covers_the_receptive_field(speech_setup):
    from mbexwn_vocoder_b200.long_form import main_context_frames, subnet_reach_frames
    hp, plan, w = speech_setup
    wn = plan.wavenet
    ctx = main_context_frames(plan)
    assert ctx * wn.steps_per_frame >= sum(d * (wn.k - 1) // 2 for d in wn.dilations) + plan.pqmf_q
    assert ctx >= subnet_reach_frames(plan.ps_ops) and ctx * plan.hop >= plan.stft_win
    assert subnet_reach_frames(plan.pp_ops) >= 2


def test_create_validates_the_wavenet_block_table():
    """mbexwn_create checks a multi-block description before it touches CUDA: block count, per-block fields, the rate chain
    steps_per_frame x prod(up) = hop / subbands and block rate = cond_conv_up x cond_lin_up (custom_pulsed_generator.py:344, :469)."""
    from mbexwn_vocoder_b200 import get_config_file
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.engine import make_config
    from mbexwn_vocoder_b200.plan import build_plan
    hp = read_config(get_config_file("SPEECH"))
    hp["mbexwn_config"].update({"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                                "pp_mod_subnet_channel_factors": [0.5, 0.25]})
    plan = build_plan(hp)
    lib = _cabi.load()

    def create(mutate):
        cfg = make_config(plan)
        mutate(cfg)
        h = ctypes.c_void_p()
        rc = lib.mbexwn_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc == _cabi.OK:
            lib.mbexwn_destroy(h)
        return rc

    assert create(lambda c: None) in (_cabi.OK, _cabi.ERR_CUDA)            # valid: OK on a GPU box, ERR_CUDA without a device
    assert create(lambda c: setattr(c, "wn_n_blocks", _cabi.MAX_BLOCKS + 1)) == _cabi.ERR_INVALID
    assert create(lambda c: setattr(c.wn_blocks[0], "up", 3)) == _cabi.ERR_INVALID           # 10 x 3 x 15 != 300
    assert create(lambda c: setattr(c.wn_blocks[1], "cond_conv_up", 1)) == _cabi.ERR_INVALID  # 20 rows per frame != 1 x 10
    assert create(lambda c: setattr(c.wn_blocks[0], "c", 0)) == _cabi.ERR_INVALID
    assert create(lambda c: setattr(c.wn_blocks[0], "up_name", b"")) == _cabi.ERR_INVALID    # up = 2 needs its conv
    assert create(lambda c: setattr(c, "abi_version", _cabi.ABI_VERSION - 1)) == _cabi.ERR_INVALID
