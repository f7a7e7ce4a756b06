"""The C-ABI fails loudly: every misuse returns its status code with a message, nothing is silently ignored."""
import ctypes as C

import numpy as np
import pytest
import torch

from mbexwn_vocoder_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(speech_setup):
    from mbexwn_vocoder_b200.engine import Engine
    hp, plan, w = speech_setup
    return Engine(plan, w, device=0)


def _err(eng):
    return eng.lib.mbexwn_last_error(eng._handle).decode()


def test_create_rejects_bad_configs(speech_setup):
    from mbexwn_vocoder_b200.engine import make_config
    hp, plan, w = speech_setup
    lib = _cabi.load()
    h = C.c_void_p()
    cfg = make_config(plan)
    cfg.abi_version = 1
    assert lib.mbexwn_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID          # stale ABI
    cfg = make_config(plan)
    cfg.hop = 301
    assert lib.mbexwn_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID          # rate algebra (custom_pulsed_generator.py:344)
    cfg = make_config(plan)
    cfg.stft_win = 1024
    assert lib.mbexwn_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_UNSUPPORTED      # win != 4 hop
    cfg = make_config(plan)
    cfg.ps_mode = 5
    assert lib.mbexwn_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID
    assert lib.mbexwn_create(None, C.byref(h)) == _cabi.ERR_INVALID


def test_forward_argument_errors(eng):
    pb = eng.prepare([12, 5], "f16f8", True)
    pb.load([np.zeros((12, 80), np.float32), np.zeros((5, 80), np.float32)],
            [np.zeros(12 * 20, np.float32), np.zeros(5 * 20, np.float32)])
    lib, h = eng.lib, eng._handle
    stream = torch.cuda.current_stream().cuda_stream
    ws, nb = pb.workspace.data_ptr(), pb.workspace.numel()
    assert lib.mbexwn_forward(h, C.byref(pb.batch), 9, ws, nb, stream) == _cabi.ERR_INVALID and "precision" in _err(eng)
    assert lib.mbexwn_forward(h, C.byref(pb.batch), pb.prec, ws, 1024, stream) == _cabi.ERR_INVALID and "workspace" in _err(eng)
    assert lib.mbexwn_forward(h, None, pb.prec, ws, nb, stream) == _cabi.ERR_INVALID
    bad = _cabi.Batch.from_buffer_copy(pb.batch)
    bad.mel = None
    assert lib.mbexwn_forward(h, C.byref(bad), pb.prec, ws, nb, stream) == _cabi.ERR_INVALID and "null" in _err(eng)
    bad = _cabi.Batch.from_buffer_copy(pb.batch)
    bad.n_utt = 0
    assert lib.mbexwn_forward(h, C.byref(bad), pb.prec, ws, nb, stream) == _cabi.ERR_INVALID and "empty" in _err(eng)
    assert lib.mbexwn_forward_host(h, C.byref(pb.batch), pb.prec, None, None, pb.out_host.data_ptr(), ws, nb, stream) == _cabi.ERR_INVALID
    assert lib.mbexwn_forward_host_begin(h, 2, C.byref(pb.batch), pb.prec, pb.mel_host.data_ptr(), None, pb.out_host.data_ptr(),
                                         ws, nb, stream) == _cabi.ERR_INVALID          # slot must be 0 or 1
    assert lib.mbexwn_forward_host_wait(h, 0) == _cabi.OK                              # nothing in flight: returns at once
    pb.run_host()                                                                      # and the handle still works
    assert np.isfinite(pb.out_host.numpy()).all()


def test_taps_options_tensors(eng):
    lib, h = eng.lib, eng._handle
    off, nb = C.c_size_t(), C.c_size_t()
    assert lib.mbexwn_tap(h, b"no_such_tap", 30, 4, 3, C.byref(off), C.byref(nb)) == _cabi.ERR_MISSING and "no_such_tap" in _err(eng)
    assert lib.mbexwn_tap(h, b"F0", 30, 4, 3, C.byref(off), C.byref(nb)) == _cabi.OK and nb.value == 30 * 100 * 4
    assert lib.mbexwn_set_option(h, b"no_such_option", 1) == _cabi.ERR_INVALID and "no_such_option" in _err(eng)
    assert lib.mbexwn_set_tensor(h, b"x", None, 16) == _cabi.ERR_INVALID
    assert lib.mbexwn_workspace_bytes(h, 0, 0, 3) == 0
    ms = (C.c_float * _cabi.N_STAGES)()
    eng.set_option("stage_timing", 0)
    pb = eng.prepare([8], "f16f8", False)
    pb.load([np.zeros((8, 80), np.float32)])
    pb.run_host()
    assert lib.mbexwn_stage_ms(h, ms) != _cabi.OK                                      # no timing was recorded


def test_missing_tensor_is_reported_by_name(speech_setup):
    """A handle that never received its weights names the first tensor it misses."""
    from mbexwn_vocoder_b200.engine import make_config
    hp, plan, w = speech_setup
    lib = _cabi.load()
    h = C.c_void_p()
    assert lib.mbexwn_create(C.byref(make_config(plan)), C.byref(h)) == _cabi.OK
    from mbexwn_vocoder_b200.sched import make_layout
    L = make_layout([6], 2, plan.pulse_per_frame)
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(getattr(L, k)).to(dev) for k in ("frame_utt", "utt_begin", "utt_end", "chunk_first")}
    mel = torch.zeros(L.n_frames, 80, device=dev)
    out = torch.zeros(L.n_frames * 300, device=dev)
    b = _cabi.Batch()
    b.n_utt, b.n_frames, b.n_chunks = 1, L.n_frames, L.n_chunks
    b.frame_utt, b.utt_begin, b.utt_end, b.chunk_first = (t[k].data_ptr() for k in ("frame_utt", "utt_begin", "utt_end", "chunk_first"))
    b.mel, b.out = mel.data_ptr(), out.data_ptr()
    nbytes = lib.mbexwn_workspace_bytes(h, L.n_frames, L.n_chunks, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = lib.mbexwn_forward(h, C.byref(b), 0, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    msg = lib.mbexwn_last_error(h).decode()
    assert rc == _cabi.ERR_MISSING and "tensor not registered: PulsPar_" in msg
    lib.mbexwn_destroy(h)
