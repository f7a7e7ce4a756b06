"""Device engine: owns the PyTorch-allocated buffers and drives libmbexwn_b200.so through the C-ABI.

PyTorch is used for device memory, streams and pinned host memory only; every arithmetic op of the forward
path runs in the hand-written CUDA kernels behind ``mbexwn_forward`` (include/mbexwn.h).  There is no CPU
fallback: constructing an Engine without a CUDA device or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import sys
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi
from . import weights as W
from .plan import ACT_PRELU, ModelPlan, Op
from .sched import FrameGridLayout, make_layout

_TAP_DTYPES = {"index": torch.int32, "lifter_index": torch.int32}


def _fill_op(dst: _cabi.Op, op: Op):
    dst.kind = 0 if op.kind == "conv" else 1
    if op.conv is not None:
        cv = op.conv
        dst.k, dst.cin, dst.cout, dst.dilation = cv.k, cv.cin, cv.cout, cv.dilation
        dst.pad_l, dst.pad_r, dst.pad_mode, dst.subpixel = cv.pad_l, cv.pad_r, cv.pad_mode, cv.subpixel
        dst.name = cv.name.encode()
    dst.up = op.up
    dst.act = op.act
    dst.act_channels = op.act_channels
    dst.rate_in, dst.rate_out, dst.ch_out = op.rate_in, op.rate_out, op.ch_out
    dst.act_name = (op.act_name or "").encode()


def make_config(plan: ModelPlan) -> _cabi.Config:
    c = _cabi.Config()
    c.abi_version = _cabi.ABI_VERSION
    c.sample_rate, c.hop, c.mel_channels = plan.sample_rate, plan.hop, plan.mel_channels
    c.pulse_per_frame, c.steps_per_frame = plan.pulse_per_frame, plan.steps_per_frame
    c.pulse_channels, c.subbands = plan.pulse_channels, plan.subbands
    c.pulse_rate, c.f0_min, c.f0_max = plan.pulse_rate, plan.f0_min, plan.f0_max
    c.f0_span = plan.f0_max - plan.f0_min
    c.noise_sigma, c.leaky_alpha = plan.noise_sigma, plan.alpha
    if len(plan.pp_ops) > _cabi.MAX_OPS or len(plan.ps_ops) > _cabi.MAX_OPS:
        raise NotImplementedError("sub-net too deep for the C-ABI op table")
    c.n_pp_ops, c.n_ps_ops = len(plan.pp_ops), len(plan.ps_ops)
    for i, op in enumerate(plan.pp_ops):
        _fill_op(c.pp_ops[i], op)
    for i, op in enumerate(plan.ps_ops):
        _fill_op(c.ps_ops[i], op)
    wn = plan.wavenet
    if wn.n_layers > _cabi.MAX_LAYERS:
        raise NotImplementedError("too many WaveNet layers for the C-ABI")
    c.wn_c, c.wn_cin, c.wn_cout, c.wn_layers, c.wn_k, c.wn_gate = wn.c, wn.c_in, wn.c_out, wn.n_layers, wn.k, wn.gate
    c.wn_cond_k, c.wn_cond_conv_up, c.wn_cond_lin_up = wn.cond_k, wn.cond_conv_up, wn.cond_lin_up
    for i, d in enumerate(wn.dilations):
        c.wn_dilations[i] = d
    c.wn_name = (wn.name + "_WNBlock_WN").encode()
    blocks = plan.blocks
    if len(blocks) > 1 or blocks[0].up > 1:                # pp_waveNetBlocks beyond the single WaveNetAE of the released models
        if len(blocks) > _cabi.MAX_BLOCKS:
            raise NotImplementedError("too many WaveNet blocks for the C-ABI")
        c.wn_n_blocks = len(blocks)
        for i, b in enumerate(blocks):
            dst = c.wn_blocks[i]
            dst.c, dst.cond_conv_up, dst.up = b.c, b.cond_conv_up, b.up
            dst.name = (b.name + "_WNBlock_WN").encode()
            dst.up_name = b.up_name.encode() if b.up > 1 else b""
    c.post_name = plan.post_name.encode()
    c.n_ceps, c.stft_win, c.fft_size = plan.n_ceps, plan.stft_win, plan.fft_size
    c.n_lifters = 0 if plan.lifters is None else int(plan.lifters.shape[0])
    c.n_smooth = int(plan.f0_smooth.shape[0])
    c.filter_max_log_range = float(plan.filter_max_log_range or 0.0)
    wt = plan.wavetables
    c.wt_n_period, c.wt_n_tables = wt.n_period, int(wt.tables.shape[1])
    c.wt_nominal_f0 = float(np.float32(wt.nominal_f0))
    c.wt_min_transposition, c.wt_max_transposition = float(wt.min_transposition), float(wt.max_transposition)
    c.wt_grid_norm = float(wt.grid_norm)
    c.cumsum_chunk = 1000
    c.pqmf_q, c.pqmf_back = plan.pqmf_q, plan.pqmf_back
    c.halo_frames = engine_halo(plan)
    c.ps_mode, c.ps_preserve_energy = plan.ps_mode, int(plan.ps_preserve_energy)
    c.wt_subharm = plan.subharm
    c.wn_causal = int(wn.causal)
    c.pulse_pqmf_taps = int(plan.pulse_pqmf_cfg["taps"]) if plan.pulse_pqmf_cfg is not None else 0
    if plan.norm is not None:
        nm = plan.norm
        c.norm_enable, c.norm_iters, c.norm_win, c.norm_smooth_win = 1, nm.iters, nm.win, nm.smooth_win
        c.norm_proj_cols, c.norm_use_max_limit = nm.proj_cols, int(nm.use_max_limit)
        c.norm_fact, c.norm_floor, c.norm_compress_exp, c.norm_proj_scale = nm.norm_fact, nm.floor, nm.compress_exp, nm.proj_scale
        c.norm_lin_scale, c.norm_lin_off, c.norm_mel_scale = nm.lin_scale, nm.lin_off, nm.mel_scale
    return c


def engine_halo(plan: ModelPlan) -> int:
    """Guard frames between utterances: covers the widest dilated tap and the PQMF polyphase reach."""
    sub = plan.sub_per_frame or plan.steps_per_frame        # rows per frame of the sub-band signals
    halo = max(1, plan.max_halo_frames, -(-max(plan.pqmf_back, plan.pqmf_q - 1 - plan.pqmf_back) // sub))
    # the tensor-core sub-net convs read their SYMMETRIC / EDGE pad rows from the guard rows next to each utterance:
    # the guard gap must hold the right pad of one utterance and the left pad of the next without overlap
    for ops in (plan.pp_ops, plan.ps_ops):
        for op in ops:
            if op.kind == "conv" and op.conv.pad_mode != 0:
                halo = max(halo, -(-(op.conv.pad_l + op.conv.pad_r) // op.rate_in))
            elif op.kind == "conv":                          # zero padding is read from the guard rows
                halo = max(halo, -(-max(op.conv.pad_l, op.conv.pad_r) // op.rate_in))
    wn = plan.wavenet
    halo = max(halo, wn.cond_k - 1 if wn.causal else (wn.cond_k - 1) // 2)     # conditioning conv at mel rate
    return halo


class Engine:
    PB_CACHE_SIZE = 6

    def __init__(self, plan: ModelPlan, weights: Dict[str, np.ndarray], device: Union[int, str, torch.device] = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("mbexwn_vocoder_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _cabi.load()
        self.plan = plan
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        W.check(plan, weights)
        self.cfg = make_config(plan)
        self.halo = self.cfg.halo_frames
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.mbexwn_create(C.byref(self.cfg), C.byref(self._handle))
        if rc != _cabi.OK:
            raise RuntimeError(f"mbexwn_create failed ({rc})")
        self._tensors: Dict[str, torch.Tensor] = {}
        self._workspace: Optional[torch.Tensor] = None
        self._pb_cache: "OrderedDict" = OrderedDict()
        self._upload(weights)
        self.last_layout: Optional[FrameGridLayout] = None
        self._last = None
        self.use_graphs = True
        self._aux_stream = None

    # ---- weights / constants ----------------------------------------------------------------------
    def _register(self, name: str, array: np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(array)).to(self.device)
        self._tensors[name] = t
        _cabi.check(self.lib, self._handle,
                    self.lib.mbexwn_set_tensor(self._handle, name.encode(), t.data_ptr(), t.numel() * t.element_size()),
                    f"set_tensor({name})")

    def _upload(self, weights: Dict[str, np.ndarray]):
        plan = self.plan
        for layer in plan.conv_layers():
            w, b = W.folded(weights, layer.name)
            self._register(f"{layer.name}/W", w)
            self._register(f"{layer.name}/b", b)
            if b.size == 1:                                  # lets the fused sub-net tail take the bias as an argument
                _cabi.check(self.lib, self._handle,
                            self.lib.mbexwn_set_scalar(self._handle, f"{layer.name}/b".encode(), float(b[0])), "set_scalar")
        for ops in (plan.pp_ops, plan.ps_ops):
            for op in ops:
                if op.act == ACT_PRELU and op.act_name:
                    self._register(f"{op.act_name}/alpha", weights[f"{op.act_name}/alpha"].astype(np.float32))
        self._register("wavetable", plan.wavetables.tables)
        self._register("pqmf_poly", plan.pqmf_poly)
        self._register("window", plan.window)
        self._register("inv_window", plan.inv_window)
        n = plan.fft_size
        ang = -2.0 * np.pi * np.arange(n // 2) / n
        self._register("twiddle", np.stack((np.cos(ang), np.sin(ang)), axis=1).astype(np.float32))
        self._register("f0_smooth", plan.f0_smooth)
        if plan.lifters is not None:
            self._register("lifters", plan.lifters)
            self._register("lifter_grid", plan.lifter_log10f0)
        if plan.pulse_pqmf_ana is not None:
            self._register("pulse_pqmf", plan.pulse_pqmf_ana)
        if plan.norm is not None:
            self._register("norm/proj", plan.norm.proj)
            self._register("norm/smooth_win", plan.norm.smooth_window)
            self._register("norm/gwin", plan.norm.gwin)
        self._upload_tc(weights)

    def _register_torch(self, name: str, t: torch.Tensor):
        t = t.contiguous().to(self.device)
        self._tensors[name] = t
        _cabi.check(self.lib, self._handle,
                    self.lib.mbexwn_set_tensor(self._handle, name.encode(), t.data_ptr(), t.numel() * t.element_size()),
                    f"set_tensor({name})")

    def _upload_tc(self, weights: Dict[str, np.ndarray]):
        from .tc_pack import pack_subnet_weights, pack_tc8_weights, pack_tc_weights
        for name, t in pack_tc_weights(self.plan, weights).items():
            self._register_torch(name, t)
        tc8, shifts = pack_tc8_weights(self.plan, weights)
        for name, t in tc8.items():
            self._register_torch(name, t)
        for opt, v in shifts.items():
            self.set_option(opt, v)
        self.tc8_shifts = shifts
        for name, t in pack_subnet_weights(self.plan, weights).items():
            self._register_torch(name, t)

    def tc_gemm(self, a: torch.Tensor, b: torch.Tensor, kblocks: np.ndarray) -> torch.Tensor:
        """Unit-test hook for the tap-GEMM kernel: a (rows, a_cols) bf16, b (n, b_cols) bf16 on the device."""
        assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
        kb = np.ascontiguousarray(kblocks, dtype=np.int32)
        out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=self.device)
        rc = self.lib.mbexwn_k_tc_gemm(self._handle, a.data_ptr(), a.shape[0], a.shape[1], b.data_ptr(), b.shape[0],
                                       b.shape[1], kb.ctypes.data, kb.shape[0], out.data_ptr(),
                                       torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(self.lib, self._handle, rc, "mbexwn_k_tc_gemm")
        return out

    def tc_gemm_f16f8(self, a: torch.Tensor, b: torch.Tensor, kblocks: np.ndarray) -> torch.Tensor:
        """Unit-test hook for the split-precision tap-GEMM: a (rows, 4 * a_cpad) uint8, b (n, 4 * b_k) uint8 rows in the
        [fp16 | e4m3 | e4m3] plane layout of include/mbexwn.h."""
        assert a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.is_cuda and b.is_cuda
        kb = np.ascontiguousarray(kblocks, dtype=np.int32)
        out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=self.device)
        rc = self.lib.mbexwn_k_tc_gemm_f16f8(self._handle, a.data_ptr(), a.shape[0], a.shape[1] // 4, b.data_ptr(),
                                             b.shape[0], b.shape[1] // 4, kb.ctypes.data, kb.shape[0], out.data_ptr(),
                                             torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(self.lib, self._handle, rc, "mbexwn_k_tc_gemm_f16f8")
        return out

    def copy_stream(self):
        """Side stream for large device->host result copies (kept apart from aux_stream, whose small geometry uploads gate the
        next forward)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        return self._copy_stream

    def aux_stream(self):
        """Side stream for small asynchronous uploads (batch geometry) that must not queue behind a running forward."""
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream(self.device)
        return self._aux_stream

    def range_status(self, reset: bool = True) -> int:
        """Sticky range-guard word of the f16f8 path (include/mbexwn.h: mbexwn_range_status): bit 0 = a residual-stream value
        left the range of the e4m3 hi8 plane (|x| > 448), bit 1 = beyond 60000 (fp16).  Call after the forward's stream has
        been synchronised (forward / run_host / wait_host do that)."""
        flags = C.c_int32(0)
        _cabi.check(self.lib, self._handle, self.lib.mbexwn_range_status(self._handle, C.byref(flags), 1 if reset else 0), "range_status")
        return int(flags.value)

    def set_option(self, name: str, value: int):
        _cabi.check(self.lib, self._handle, self.lib.mbexwn_set_option(self._handle, name.encode(), value), "set_option")

    def get_info(self, name: str) -> int:
        """mbexwn_get_info: what the last forward did ("tc_last_fused", "tc_last_cluster", "tc_max_quads")."""
        import ctypes
        v = ctypes.c_int32(0)
        _cabi.check(self.lib, self._handle, self.lib.mbexwn_get_info(self._handle, name.encode(), ctypes.byref(v)), "get_info")
        return int(v.value)

    # ---- forward ----------------------------------------------------------------------------------
    def _ensure_workspace(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    def _grid_tensors(self, layout: FrameGridLayout):
        dev = self.device
        return (torch.from_numpy(layout.frame_utt).to(dev), torch.from_numpy(layout.utt_begin).to(dev),
                torch.from_numpy(layout.utt_end).to(dev), torch.from_numpy(layout.chunk_first).to(dev))

    def prepare(self, lengths: Sequence[int], precision: str = "fp32", with_noise: bool = True,
                with_f0: bool = False, with_carry: bool = False, capacity_frames: int = 0,
                capacity_utts: int = 0) -> "PreparedBatch":
        """Allocate everything one batch geometry needs (device grid, staging, pinned host buffers, workspace); with a
        capacity, `PreparedBatch.rebind` later switches geometry without allocating."""
        return PreparedBatch(self, lengths, precision, with_noise, with_f0, with_carry, capacity_frames, capacity_utts)

    def prepare_cached(self, lengths: Sequence[int], precision: str = "fp32", with_noise: bool = True,
                       with_f0: bool = False, slot: int = 0, with_carry: bool = False) -> "PreparedBatch":
        """`prepare` with a small LRU of batch geometries: repeated calls of one shape (a client sending utterance after
        utterance, the windows of long-form synthesis) skip the pinned-memory and device allocations, which cost far more
        than a short forward.  `slot` separates buffer sets that are in flight at the same time (synth_stream)."""
        key = (tuple(int(t) for t in lengths), precision, bool(with_noise), bool(with_f0), int(slot), bool(with_carry))
        pb = self._pb_cache.pop(key, None)
        if pb is None or pb.workspace.data_ptr() != (self._workspace.data_ptr() if self._workspace is not None else 0):
            pb = PreparedBatch(self, lengths, precision, with_noise, with_f0, with_carry)
            # a larger workspace may have replaced the one cached batches point to: drop those
            live = self._workspace.data_ptr()
            for k in [k for k, v in self._pb_cache.items() if v.workspace.data_ptr() != live]:
                del self._pb_cache[k]
        self._pb_cache[key] = pb
        while len(self._pb_cache) > self.PB_CACHE_SIZE:
            self._pb_cache.pop(next(iter(self._pb_cache)))
        pb.batch.utt_ids = None
        return pb

    # batches of at most this many padded frames replay a CUDA graph of their forward (captured per cached batch geometry):
    # the launch-bound case of BASELINE.json configs[0] (one 5 s utterance: ~25 kernel launches of a few microseconds each)
    GRAPH_MAX_FRAMES = 1024

    def forward(self, mels: Sequence[np.ndarray], noise: Optional[Sequence[np.ndarray]] = None,
                f0: Optional[Sequence[np.ndarray]] = None, precision: str = "fp32", seed: int = 0,
                taps: Sequence[str] = (), utt_ids: Optional[Sequence[int]] = None):
        """mels: list of (T_u, n_mel) float32.  Returns (list of (T_u*hop,) waveforms, {tap: list of arrays}).

        utt_ids: global utterance ids keying the in-kernel noise stream (default: position in this batch)."""
        pb = self.prepare_cached([m.shape[0] for m in mels], precision, noise is not None, f0 is not None)
        pb.load(mels, noise, f0)
        if utt_ids is not None:
            pb.set_utt_ids(utt_ids)
        if self.use_graphs and not taps and utt_ids is None and pb.layout.n_frames <= self.GRAPH_MAX_FRAMES:
            pb.run_graph(seed)
        else:
            pb.run_host(seed)
        out = pb.waveforms_copy()
        return out, {t: pb.tap(t) for t in taps}

    def close(self):
        if self._handle:
            self.lib.mbexwn_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PreparedBatch:
    """One batch geometry bound to its buffers; `run_host` is the reference-facing call (host in, host out)."""

    def __init__(self, eng: Engine, lengths: Sequence[int], precision: str, with_noise: bool, with_f0: bool,
                 with_carry: bool = False, capacity_frames: int = 0, capacity_utts: int = 0):
        """`capacity_frames` / `capacity_utts` > 0: allocate for that many padded frames / utterances so that `rebind` can
        switch to any other geometry within the capacity without allocating (ragged serving, config 4)."""
        plan = eng.plan
        self.eng = eng
        self.prec = _cabi.PRECISIONS[precision]
        self.layout = make_layout(lengths, eng.halo, plan.pulse_per_frame, eng.cfg.cumsum_chunk)
        L = self.layout
        dev = eng.device
        F = max(L.n_frames, int(capacity_frames))
        U = max(L.n_utt, int(capacity_utts))
        self.cap_frames, self.cap_utts = F, U
        # worst case of ceil(T_u * pulse_per_frame / chunk) summed over the utterances
        self.cap_chunks = max(L.n_chunks, -(-F * plan.pulse_per_frame // eng.cfg.cumsum_chunk) + U)
        # batch geometry (frame -> utterance map, utterance bounds, chunk table, global utterance ids): ONE device buffer fed
        # from ONE pinned staging buffer by an asynchronous copy on the engine's auxiliary stream.  A pageable copy on the
        # compute stream would queue behind the forward of the other buffer set and stall the host for a whole batch
        # (this serialised host preparation and GPU work in the many-utterance path).
        n_meta = F + 4 * U + 8
        self.meta_host = torch.zeros(n_meta, dtype=torch.int32).pin_memory()
        # torch.empty, not zeros: a fill kernel on the compute stream could run AFTER the first upload on the auxiliary stream and
        # wipe the geometry (every frame a guard frame: an all-zero waveform).  The block may also have had earlier users whose
        # work is still queued on the compute stream: the first upload waits for the event recorded here.
        self.meta_dev = torch.empty(n_meta, dtype=torch.int32, device=dev)
        self._alloc_event = torch.cuda.Event()
        self._alloc_event.record(torch.cuda.current_stream(dev))
        self.frame_utt = self.meta_dev[:F]
        self.utt_begin = self.meta_dev[F:F + U]
        self.utt_end = self.meta_dev[F + U:F + 2 * U]
        self.chunk_first = self.meta_dev[F + 2 * U:F + 3 * U + 1]
        self._utt_ids_dev = self.meta_dev[F + 3 * U + 4:F + 4 * U + 4]
        self._meta_dirty = False
        self._meta_event = None
        self.mel_host = torch.zeros(F, plan.mel_channels, dtype=torch.float32).pin_memory()
        self.out_host = torch.zeros(F * plan.hop, dtype=torch.float32).pin_memory()
        self.mel_dev = torch.zeros(F, plan.mel_channels, dtype=torch.float32, device=dev)
        self.out_dev = torch.zeros(F * plan.hop, dtype=torch.float32, device=dev)
        self.noise_host = self.noise_dev = self.f0_dev = None
        if with_noise:
            self.noise_host = torch.zeros(F * plan.steps_per_frame, dtype=torch.float32).pin_memory()
            self.noise_dev = torch.zeros(F * plan.steps_per_frame, dtype=torch.float32, device=dev)
        if with_f0:
            self.f0_dev = torch.zeros(F * plan.pulse_per_frame, dtype=torch.float32, device=dev)
        self.ws_bytes = int(eng.lib.mbexwn_workspace_bytes(eng._handle, F, self.cap_chunks, self.prec))
        self.workspace = eng._ensure_workspace(self.ws_bytes)
        self.utt_ids = None
        self._graphs = {}                                     # seed -> CUDA graph of [H2D, forward, D2H] (run_graph)
        self.batch = _cabi.Batch()
        b = self.batch
        b.frame_utt, b.utt_begin, b.utt_end = self.frame_utt.data_ptr(), self.utt_begin.data_ptr(), self.utt_end.data_ptr()
        b.chunk_first = self.chunk_first.data_ptr()
        b.mel = self.mel_dev.data_ptr()
        b.noise = self.noise_dev.data_ptr() if with_noise else None
        b.f0_override = self.f0_dev.data_ptr() if with_f0 else None
        b.out = self.out_dev.data_ptr()
        self.carry_dev = None
        if with_carry:                                        # windows of a longer signal (long_form.py)
            self.carry_dev = torch.zeros(U, dtype=torch.float32, device=dev)
            b.phase_carry = self.carry_dev.data_ptr()
        self._bind(L)
        # The buffers above were zero-filled by kernels queued on the compute stream, possibly behind a forward that is still
        # running; the pipelined host path copies into them from its own copy streams, which do not wait for the compute
        # stream.  Creating a buffer set is rare (prepare_cached) and slow anyway (pinned allocations): finish the fills here.
        torch.cuda.current_stream(dev).synchronize()

    def _bind(self, L: FrameGridLayout):
        self.layout = L
        F, U = self.cap_frames, self.cap_utts
        mh = self.meta_host.numpy()
        mh[:L.n_frames] = L.frame_utt
        mh[F:F + L.n_utt] = L.utt_begin
        mh[F + U:F + U + L.n_utt] = L.utt_end
        mh[F + 2 * U:F + 2 * U + L.n_utt + 1] = L.chunk_first
        self._meta_dirty = True
        b = self.batch
        b.n_utt, b.n_frames, b.n_chunks = L.n_utt, L.n_frames, L.n_chunks

    def rebind(self, lengths: Sequence[int]):
        """Switch to another batch geometry inside the allocated capacity.  The caller must have drained the previous use
        (run_host returned / wait_host).  Guard rows of the host staging buffers are cleared."""
        plan = self.eng.plan
        L = make_layout(lengths, self.eng.halo, plan.pulse_per_frame, self.eng.cfg.cumsum_chunk)
        if L.n_frames > self.cap_frames or L.n_utt > self.cap_utts or L.n_chunks > self.cap_chunks:
            raise RuntimeError(f"batch of {L.n_frames} padded frames / {L.n_utt} utterances exceeds the prepared capacity "
                               f"({self.cap_frames} / {self.cap_utts})")
        # only the guard rows need clearing: every utterance row is overwritten by the next load (clearing the whole grid cost
        # more host time than scattering the mels into it -- 10 MB per 32k-frame batch)
        mh = self.mel_host.numpy()
        nh = self.noise_host.numpy() if self.noise_host is not None else None
        spf = plan.steps_per_frame
        prev = 0
        for u in range(L.n_utt + 1):
            nxt = int(L.utt_begin[u]) if u < L.n_utt else L.n_frames
            if nxt > prev:
                mh[prev:nxt] = 0.0
                if nh is not None:
                    nh[prev * spf:nxt * spf] = 0.0
            if u < L.n_utt:
                prev = int(L.utt_end[u])
        self._bind(L)
        self.batch.utt_ids = None
        return self

    def set_utt_ids(self, utt_ids: Sequence[int]):
        ids = np.asarray(utt_ids, dtype=np.int32)
        assert ids.shape == (self.layout.n_utt,)
        F, U = self.cap_frames, self.cap_utts
        self.meta_host.numpy()[F + 3 * U + 4:F + 3 * U + 4 + ids.size] = ids
        self.utt_ids = self._utt_ids_dev
        self.batch.utt_ids = self._utt_ids_dev.data_ptr()
        self._meta_dirty = True

    def _flush_meta(self):
        """Upload the batch geometry if it changed: asynchronous copy on the auxiliary stream, the compute stream waits on the
        device (the host never blocks).  The previous user of these device arrays was this buffer set's own previous forward,
        which the caller has drained (wait_host / run_host) before re-binding."""
        if not self._meta_dirty:
            return
        eng = self.eng
        with torch.cuda.device(eng.device):
            aux = eng.aux_stream()
            if self._alloc_event is not None:
                aux.wait_event(self._alloc_event)
                self._alloc_event = None
            with torch.cuda.stream(aux):
                self.meta_dev.copy_(self.meta_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(aux)
            torch.cuda.current_stream(eng.device).wait_event(ev)
            self._meta_event = ev
        self._meta_dirty = False

    # bytes moved per run_host call
    @property
    def h2d_bytes(self) -> int:
        plan, F = self.eng.plan, self.layout.n_frames
        return F * plan.mel_channels * 4 + (F * plan.steps_per_frame * 4 if self.noise_host is not None else 0)

    @property
    def d2h_bytes(self) -> int:
        return self.layout.n_frames * self.eng.plan.hop * 4

    def load(self, mels, noise=None, f0=None, carry=None):
        L, plan = self.layout, self.eng.plan
        self.layout.scatter([np.asarray(m, dtype=np.float32) for m in mels], 1, self.mel_host.numpy())
        if noise is not None:
            L.scatter([np.asarray(z, dtype=np.float32).reshape(-1) for z in noise], plan.steps_per_frame,
                      self.noise_host.numpy())
        if f0 is not None:
            buf = np.zeros(L.n_frames * plan.pulse_per_frame, dtype=np.float32)
            L.scatter([np.asarray(x, dtype=np.float32).reshape(-1) for x in f0], plan.pulse_per_frame, buf)
            self.f0_dev[:buf.size].copy_(torch.from_numpy(buf))
        if carry is not None:
            carry = np.asarray(carry, dtype=np.float32)
            self.carry_dev[:carry.size].copy_(torch.from_numpy(carry))

    def _stream(self):
        return torch.cuda.current_stream(self.eng.device).cuda_stream

    def run_host(self, seed: int = 0):
        """Host mel in -> host waveform out through mbexwn_forward_host (copies inside the call)."""
        eng = self.eng
        self.batch.seed = seed
        self._flush_meta()
        with torch.cuda.device(eng.device):
            rc = eng.lib.mbexwn_forward_host(
                eng._handle, C.byref(self.batch), self.prec, self.mel_host.data_ptr(),
                self.noise_host.data_ptr() if self.noise_host is not None else None, self.out_host.data_ptr(),
                self.workspace.data_ptr(), self.workspace.numel(), self._stream())
        _cabi.check(eng.lib, eng._handle, rc, "mbexwn_forward_host")

    def begin_host(self, slot: int, seed: int = 0):
        """Pipelined host forward (mbexwn_forward_host_begin): returns once everything is enqueued.  Use two PreparedBatch
        objects of one engine alternately as slot 0 / 1 and call wait_host(slot) before touching out_host / mel_host."""
        eng = self.eng
        self.batch.seed = seed
        self._flush_meta()
        with torch.cuda.device(eng.device):
            rc = eng.lib.mbexwn_forward_host_begin(
                eng._handle, slot, C.byref(self.batch), self.prec, self.mel_host.data_ptr(),
                self.noise_host.data_ptr() if self.noise_host is not None else None, self.out_host.data_ptr(),
                self.workspace.data_ptr(), self.workspace.numel(), self._stream())
        _cabi.check(eng.lib, eng._handle, rc, "mbexwn_forward_host_begin")

    def wait_host(self, slot: int):
        _cabi.check(self.eng.lib, self.eng._handle, self.eng.lib.mbexwn_forward_host_wait(self.eng._handle, slot),
                    "mbexwn_forward_host_wait")

    def run_graph(self, seed: int = 0):
        """Host mel in -> host waveform out through a CUDA graph of [H2D, forward, D2H] captured once per (geometry, seed): one
        graph launch instead of ~25 kernel launches and as many cuTensorMapEncode calls.  Falls back to run_host if the capture
        fails (and stays there for this batch geometry)."""
        eng = self.eng
        key = int(seed)
        g = self._graphs.get(key) if self._graphs is not None else None
        if g is None and self._graphs is not None:
            try:
                n = self.layout.n_frames * eng.plan.hop
                with torch.cuda.device(eng.device):
                    side = torch.cuda.Stream(eng.device)
                    side.wait_stream(torch.cuda.current_stream(eng.device))
                    with torch.cuda.stream(side):               # warm-up outside the capture: one-time initialisations
                        self.upload()
                        self.run_device(seed)
                    torch.cuda.current_stream(eng.device).wait_stream(side)
                    torch.cuda.synchronize(eng.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        self.upload()
                        self.run_device(seed)
                        self.out_host[:n].copy_(self.out_dev[:n], non_blocking=True)
                if len(self._graphs) >= 4:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = g
            except Exception as e:                              # capture is an optimisation, never a requirement
                print(f"mbexwn_vocoder_b200::note::CUDA graph capture failed ({type(e).__name__}: {e}); using plain launches",
                      file=sys.stderr)
                torch.cuda.synchronize(eng.device)
                self._graphs = None
                g = None
        if g is None:
            return self.run_host(seed)
        with torch.cuda.device(eng.device):
            g.replay()
            torch.cuda.current_stream(eng.device).synchronize()

    def upload(self):
        F, spf = self.layout.n_frames, self.eng.plan.steps_per_frame
        self.mel_dev[:F].copy_(self.mel_host[:F], non_blocking=True)
        if self.noise_host is not None:
            self.noise_dev[:F * spf].copy_(self.noise_host[:F * spf], non_blocking=True)

    def run_device(self, seed: int = 0):
        """Device-resident inputs -> device output (asynchronous on the current stream)."""
        eng = self.eng
        self.batch.seed = seed
        self._flush_meta()
        with torch.cuda.device(eng.device):
            rc = eng.lib.mbexwn_forward(eng._handle, C.byref(self.batch), self.prec, self.workspace.data_ptr(),
                                        self.workspace.numel(), self._stream())
        _cabi.check(eng.lib, eng._handle, rc, "mbexwn_forward")

    def stage_ms(self) -> Dict[str, float]:
        """Device milliseconds per stage of the last run (needs eng.set_option("stage_timing", 1))."""
        buf = (C.c_float * _cabi.N_STAGES)()
        _cabi.check(self.eng.lib, self.eng._handle, self.eng.lib.mbexwn_stage_ms(self.eng._handle, buf), "stage_ms")
        return {n: float(buf[i]) for i, n in enumerate(_cabi.STAGE_NAMES)}

    def wavenet_launch_ms(self) -> Dict[str, float]:
        """Device ms of the gate / res-skip tap-GEMM launches of the last run, summed over the layers (needs
        "stage_timing" and a tensor-core precision)."""
        g, r, n = C.c_float(), C.c_float(), C.c_int32()
        _cabi.check(self.eng.lib, self.eng._handle,
                    self.eng.lib.mbexwn_wavenet_launch_ms(self.eng._handle, C.byref(g), C.byref(r), C.byref(n)),
                    "wavenet_launch_ms")
        return {"gate": float(g.value), "resskip": float(r.value), "layers": int(n.value)}

    def launches(self) -> int:
        return int(self.eng.lib.mbexwn_last_launch_count(self.eng._handle))

    def waveforms_copy(self, threads: int = 4) -> List[np.ndarray]:
        """Fresh copies of the waveforms in the pinned output grid.  Small batches: one copy per utterance; from a few MB on the
        utterances are copied into ONE new array by a few threads (NumPy copies release the GIL: a 245 MB batch takes 30 ms on one
        host thread) and returned as views of it."""
        views = self.waveforms()
        total = sum(v.size for v in views)
        if total < (1 << 20) or threads <= 1:
            return [v.copy() for v in views]
        flat = np.empty(total, dtype=np.float32)
        offs = np.concatenate(([0], np.cumsum([v.size for v in views])))
        outs = [flat[offs[i]:offs[i + 1]] for i in range(len(views))]
        import threading

        def part(k):
            for d, v in zip(outs[k::threads], views[k::threads]):
                d[...] = v
        ts = [threading.Thread(target=part, args=(k,)) for k in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return outs

    def waveforms(self, from_device: bool = False) -> List[np.ndarray]:
        n = self.layout.n_frames * self.eng.plan.hop
        buf = self.out_dev[:n].cpu().numpy() if from_device else self.out_host.numpy()[:n]
        return self.layout.gather(buf, self.eng.plan.hop)

    def tap_grid(self, name: str) -> torch.Tensor:
        eng = self.eng
        off, nb = C.c_size_t(), C.c_size_t()
        rc = eng.lib.mbexwn_tap(eng._handle, name.encode(), self.layout.n_frames, self.layout.n_chunks, self.prec,
                                C.byref(off), C.byref(nb))
        _cabi.check(eng.lib, eng._handle, rc, f"tap({name})")
        raw = self.workspace[off.value:off.value + nb.value]
        return raw.view(_TAP_DTYPES.get(name, torch.float32))

    def tap(self, name: str) -> List[np.ndarray]:
        """Per-utterance view of a stage tap: list of (T_u * rate, channels) arrays."""
        flat = self.tap_grid(name).cpu().numpy()
        F = self.layout.n_frames
        per_frame = flat.size // F
        grid = flat.reshape(F, per_frame)
        if name == "wn_out":                                   # stored with the channels padded to a multiple of 32
            c_out = self.eng.plan.wavenet.c_out
            grid = grid.reshape(F, -1, -(-c_out // 32) * 32)[:, :, :c_out].reshape(F, -1)
        return [grid[self.layout.utt_begin[u]:self.layout.utt_end[u]].reshape(-1) for u in range(self.layout.n_utt)]
