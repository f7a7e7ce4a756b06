"""Chunked long-form synthesis (BASELINE.json configs[4]: one 10-minute mel, chunked with receptive-field overlap).

The forward pass of MBExWN (custom_pulsed_generator.py:556-771) has finite memory everywhere except in the pulse phase
(tf_wavetable.py:429-492: a cumulative sum over the whole utterance).  A long mel is therefore cut into windows
[core - context, core + context) that are processed as independent "utterances" of one batch; only the core of each
window is kept.  With a context that covers the receptive field of every stage the cores are *bit-identical* to the
un-chunked result, because every kernel works per row with a fixed summation order (tests/test_gpu_long_form.py).

Three passes:
  1. F0 pass  -- the F0 sub-net over the whole signal (one utterance of one forward that stops behind the sub-net; rows are
                 independent, so this equals any chunked evaluation bit for bit) -> exact F0 track, left on the device;
  2. phase carry -- the unwrapped running sum of the wrapped 1000-sample chunk totals (float32, the reference's sequential
                    association order), one value per cumsum chunk: mbexwn_phase_carry on the device;
  3. main pass -- windows whose start is a multiple of the cumsum chunk (10 frames) with f0_override = exact F0 slice and
                  phase_carry = the running sum before the window's first chunk.  The windows are cut out of the
                  device-resident mel / noise / F0 buffers by mbexwn_gather_rows and their cores are gathered into one
                  device buffer the same way: after the upload of the mel the host only enqueues.
Windows are batched up to `max_batch_frames` per call, so throughput comes from the same batched kernels as configs 2-4;
the first window alone gives the first-chunk latency.  `host_loop=True` keeps the round-1 form (F0 track and phase carry
through the host), used by the tests as a second opinion.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .plan import ModelPlan


def subnet_reach_frames(ops) -> int:
    """Mel frames a sub-net output frame can see on either side (conv pads and interpolation neighbours)."""
    reach = 0.0
    for op in ops:
        if op.kind == "conv":
            c = op.conv
            reach += max(c.pad_l, c.pad_r, (c.k - 1) * c.dilation - min(c.pad_l, c.pad_r)) / op.rate_in
        else:
            reach += 1.0 / op.rate_in               # LinInterp reads the next low-rate row
    return int(np.ceil(reach)) + 1


def main_context_frames(plan: ModelPlan) -> int:
    """Context (frames per side) after which a window's core no longer depends on where the window was cut, given the
    exact F0: WaveNet receptive field + conditioning conv / interpolation + VTF sub-net + PQMF + STFT + lifter smoothing."""
    wn = plan.wavenet
    span = (wn.k - 1) if wn.causal else (wn.k - 1) // 2                      # causal convs look (k - 1) d rows back
    sub = plan.sub_per_frame or wn.steps_per_frame
    ctx = -(-max(plan.pqmf_back, plan.pqmf_q - 1 - plan.pqmf_back) // sub)   # PQMF rows -> frames
    for blk in plan.blocks:                                                  # receptive fields of the blocks add up
        rows = sum(d * span for d in blk.dilations) + (2 if blk.up > 1 else 0)   # + the k = 3 up-sampling conv
        ctx += -(-rows // blk.steps_per_frame)
    ctx += ((wn.cond_k - 1) if wn.causal else (wn.cond_k - 1) // 2) + 1      # conditioning conv + its x10 interpolation
    ctx += -(-(plan.stft_win // 2) // plan.hop)                              # STFT frames overlapping a sample
    ctx = max(ctx, subnet_reach_frames(plan.ps_ops) + 2,                     # VTF sub-net (per frame, no recursion)
              -(-(len(plan.f0_smooth) // 2) // plan.pulse_per_frame) + 2)    # F0 smoothing of the lifter selection
    return ctx + 2


def phase_run_before_chunks(f0: np.ndarray, pulse_rate: float, chunk: int = 1000) -> np.ndarray:
    """Unwrapped float32 running sum of the wrapped chunk totals *before* every cumsum chunk of one utterance.

    Restates the offset part of PulseWaveTable.stable_cumsum_and_wrap (tf_wavetable.py:470-486): velocities f0 / rate are
    summed sequentially in float32 inside chunks of `chunk` samples, the totals are wrapped (floor-mod 1) and accumulated
    sequentially over the chunks without wrapping.  Same association order as phase_chunk_kernel / chunk_offset_kernel.
    """
    v = np.asarray(f0, dtype=np.float32) / np.float32(pulse_rate)
    n_chunks = -(-v.size // chunk)
    pad = np.zeros(n_chunks * chunk, dtype=np.float32)
    pad[:v.size] = v
    totals = np.cumsum(pad.reshape(n_chunks, chunk), axis=1, dtype=np.float32)[:, -1]
    wrapped = totals - np.floor(totals)
    run = np.zeros(n_chunks, dtype=np.float32)
    run[1:] = np.cumsum(wrapped, dtype=np.float32)[:-1]
    return run


@dataclass
class Window:
    start: int      # first frame of the window (multiple of `align`)
    stop: int
    core0: int      # first frame of the kept core
    core1: int


def plan_windows(n_frames: int, chunk_frames: int, context: int, align: int = 1) -> List[Window]:
    """Cores tile [0, n_frames); every window is its core plus `context` frames per side, start rounded down to `align`."""
    if chunk_frames <= 0:
        raise ValueError("chunk_frames must be positive")
    out = []
    for c0 in range(0, n_frames, chunk_frames):
        c1 = min(n_frames, c0 + chunk_frames)
        s = max(0, c0 - context)
        s -= s % align
        out.append(Window(s, min(n_frames, c1 + context), c0, c1))
    return out


def _batches(windows: Sequence[Window], max_batch_frames: int) -> List[List[int]]:
    groups, cur, frames = [], [], 0
    for i, w in enumerate(windows):
        n = w.stop - w.start
        if cur and frames + n > max_batch_frames:
            groups.append(cur)
            cur, frames = [], 0
        cur.append(i)
        frames += n
    if cur:
        groups.append(cur)
    return groups


def synth_long(engine, mel: np.ndarray, noise: np.ndarray, chunk_frames: int = 400, precision: str = "f16f8",
               max_batch_frames: int = 32768, first_alone: bool = True, host_loop: bool = False) -> Tuple[np.ndarray, Dict[str, float]]:
    """mel (T, n_mel), noise (T * steps_per_frame,) standard normal -> waveform (T * hop,), timing info.

    `first_alone`: the first window is synthesised in a call of its own (first-chunk latency of a streaming client)."""
    if not host_loop:
        return _synth_long_device(engine, mel, noise, chunk_frames, precision, max_batch_frames, first_alone)
    plan: ModelPlan = engine.plan
    if plan.norm is not None:
        raise NotImplementedError("chunked synthesis with normalize_rms_from_mell is not built (the smoothed RMS spans "
                                  "window borders)")
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    T = mel.shape[0]
    ppf, spf, hop = plan.pulse_per_frame, plan.steps_per_frame, plan.hop
    noise = np.asarray(noise, dtype=np.float32).reshape(-1)
    if noise.size != T * spf:
        raise RuntimeError(f"noise must hold {T * spf} draws, got {noise.size}")
    chunk = int(engine.cfg.cumsum_chunk)
    if chunk % ppf:
        raise NotImplementedError("cumsum chunk is not a whole number of frames")
    align = chunk // ppf
    info: Dict[str, float] = {}
    t0 = time.perf_counter()

    ctx = main_context_frames(plan)
    main_wins = plan_windows(T, chunk_frames, ctx, align)
    out = np.empty(T * hop, dtype=np.float32)
    f0_ctx = subnet_reach_frames(plan.pp_ops)

    def f0_pass(lo: int, hi: int, dst: np.ndarray, cached: bool = False):
        """Exact F0 of frames [lo, hi) (receptive-field context taken from the neighbouring frames) into dst[lo*ppf:hi*ppf]."""
        a, b = max(0, lo - f0_ctx), min(T, hi + f0_ctx)
        wins = plan_windows(b - a, max(chunk_frames, 4 * f0_ctx), f0_ctx)
        engine.set_option("stop_after_f0", 1)
        try:
            for grp in _batches(wins, max_batch_frames):
                # only the small first-window geometry is kept in the engine's buffer cache: holding the large batches
                # there would keep their pinned staging memory from being recycled
                prep = engine.prepare_cached if cached else engine.prepare
                pb = prep([wins[i].stop - wins[i].start for i in grp], precision, with_noise=False)
                pb.load([mel[a + wins[i].start:a + wins[i].stop] for i in grp])
                pb.upload()
                pb.run_device()
                for i, x in zip(grp, pb.tap("F0")):
                    w = wins[i]
                    c0, c1 = max(w.core0 + a, lo), min(w.core1 + a, hi)
                    if c1 > c0:
                        dst[c0 * ppf:c1 * ppf] = x[(c0 - a - w.start) * ppf:(c1 - a - w.start) * ppf]
        finally:
            engine.set_option("stop_after_f0", 0)

    def run_windows(grp, f0_src: np.ndarray, carries, cached: bool = False):
        prep = engine.prepare_cached if cached else engine.prepare
        pb = prep([main_wins[i].stop - main_wins[i].start for i in grp], precision, with_noise=True, with_f0=True,
                  with_carry=True)
        pb.load([mel[main_wins[i].start:main_wins[i].stop] for i in grp],
                noise=[noise[main_wins[i].start * spf:main_wins[i].stop * spf] for i in grp],
                f0=[f0_src[main_wins[i].start * ppf:main_wins[i].stop * ppf] for i in grp],
                carry=carries)
        pb.run_host()
        for i, y in zip(grp, pb.waveforms()):
            w = main_wins[i]
            out[w.core0 * hop:w.core1 * hop] = y[(w.core0 - w.start) * hop:(w.core1 - w.start) * hop]

    # ---- first window on its own: it needs the F0 of its own frames only and starts with a zero phase carry, so a
    # streaming client hears it before the rest of the utterance has been looked at -------------------------------
    f0_full = np.empty(T * ppf, dtype=np.float32)
    first_done = False
    if first_alone and len(main_wins) > 1:
        w0 = main_wins[0]
        f0_pass(w0.start, w0.stop, f0_full, cached=True)
        run_windows([0], f0_full, [0.0], cached=True)
        info["first_chunk_latency_s"] = time.perf_counter() - t0
        first_done = True

    # ---- pass 1: exact F0 of the whole signal ----------------------------------------------------------------
    t_f0 = time.perf_counter()
    f0_pass(0, T, f0_full)
    info["f0_pass_s"] = time.perf_counter() - t_f0

    # ---- pass 2: phase carry per cumsum chunk ------------------------------------------------------------------
    run = phase_run_before_chunks(f0_full, plan.pulse_rate, chunk)

    # ---- pass 3: the remaining windows, batched ---------------------------------------------------------------
    rest = list(range(1 if first_done else 0, len(main_wins)))
    groups = [[rest[j] for j in g] for g in _batches([main_wins[i] for i in rest], max_batch_frames)] if rest else []
    t1 = time.perf_counter()
    for gi, grp in enumerate(groups):
        run_windows(grp, f0_full, [run[main_wins[i].start // align] for i in grp])
        if gi == 0 and not first_done:
            info["first_chunk_latency_s"] = time.perf_counter() - t0
    info["main_pass_s"] = time.perf_counter() - t1
    info["total_s"] = time.perf_counter() - t0
    info["audio_s"] = T * hop / plan.sample_rate
    info["n_windows"] = len(main_wins)
    info["context_frames"] = ctx
    return out, info


def _synth_long_device(engine, mel: np.ndarray, noise: np.ndarray, chunk_frames: int, precision: str, max_batch_frames: int,
                       first_alone: bool) -> Tuple[np.ndarray, Dict[str, float]]:
    """Chunked synthesis with the whole signal resident on the device (module docstring)."""
    import ctypes as C

    import torch

    from . import _cabi
    plan: ModelPlan = engine.plan
    if plan.norm is not None:
        raise NotImplementedError("chunked synthesis with normalize_rms_from_mell is not built (the smoothed RMS spans "
                                  "window borders)")
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    T = mel.shape[0]
    ppf, spf, hop = plan.pulse_per_frame, plan.steps_per_frame, plan.hop
    noise = np.asarray(noise, dtype=np.float32).reshape(-1)
    if noise.size != T * spf:
        raise RuntimeError(f"noise must hold {T * spf} draws, got {noise.size}")
    chunk = int(engine.cfg.cumsum_chunk)
    if chunk % ppf:
        raise NotImplementedError("cumsum chunk is not a whole number of frames")
    align = chunk // ppf
    dev = engine.device
    lib, handle = engine.lib, engine._handle
    info: Dict[str, float] = {}
    t0 = time.perf_counter()
    ctx = main_context_frames(plan)
    wins = plan_windows(T, chunk_frames, ctx, align)
    out = np.empty(T * hop, dtype=np.float32)

    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream

        # ---- first window on its own (host buffers, cached geometry): ONE plain forward over its frames plus the reach of the
        # F0 sub-net to the right, so that the F0 it computes itself is exact on every frame the core can see (frames behind
        # w0.stop are outside the core's context); zero phase carry = a plain forward
        first_done = False
        if first_alone and len(wins) > 1:
            w0 = wins[0]
            b = min(T, w0.stop + subnet_reach_frames(plan.pp_ops))
            pb0 = engine.prepare_cached([b], precision, with_noise=True)
            pb0.load([mel[:b]], noise=[noise[:b * spf]])
            pb0.run_host()
            y = pb0.waveforms()[0]
            out[w0.core0 * hop:w0.core1 * hop] = y[w0.core0 * hop:w0.core1 * hop]
            info["first_chunk_latency_s"] = time.perf_counter() - t0
            first_done = True

        # ---- whole signal to the device; pass 1: F0 of the whole signal as one utterance (stays on the device) -----------
        t_f0 = time.perf_counter()
        pbF = engine.prepare_cached([T], precision, with_noise=False)
        pbF.load([mel])
        pbF.upload()
        engine.set_option("stop_after_f0", 1)
        try:
            pbF.run_device()
        finally:
            engine.set_option("stop_after_f0", 0)
        halo = engine.halo
        f0_full = pbF.tap_grid("F0")[halo * ppf:(halo + T) * ppf]          # view into the workspace: copied before the next forward
        key = ("long_form", T)
        bufs = getattr(engine, "_long_bufs", {}).get(key)
        if bufs is None:
            n_chunks = -(-T * ppf // chunk)
            bufs = {"f0": torch.empty(T * ppf, dtype=torch.float32, device=dev),
                    "noise": torch.empty(T * spf, dtype=torch.float32, device=dev),
                    "noise_pin": torch.empty(T * spf, dtype=torch.float32).pin_memory(),
                    "run": torch.empty(n_chunks, dtype=torch.float32, device=dev),
                    "out": torch.empty(T * hop, dtype=torch.float32, device=dev),
                    "out_pin": torch.empty(T * hop, dtype=torch.float32).pin_memory()}
            engine._long_bufs = {key: bufs}
        bufs["f0"].copy_(f0_full)
        mel_full = pbF.mel_dev[halo:halo + T]                              # the whole mel is already on the device
        bufs["noise_pin"].numpy()[:] = noise
        bufs["noise"].copy_(bufs["noise_pin"], non_blocking=True)
        # ---- pass 2: phase carry before every cumsum chunk ----------------------------------------------------------
        _cabi.check(lib, handle, lib.mbexwn_phase_carry(handle, bufs["f0"].data_ptr(), T * ppf, bufs["run"].data_ptr(), stream),
                    "mbexwn_phase_carry")
        info["f0_pass_s"] = time.perf_counter() - t_f0

        # ---- pass 3: the windows, batched; cut out and gathered back on the device -----------------------------------
        t1 = time.perf_counter()
        rest = list(range(1 if first_done else 0, len(wins)))
        groups = [[rest[j] for j in g] for g in _batches([wins[i] for i in rest], max_batch_frames)] if rest else []

        def gather(src, dst, row_elems, seg_dev, n_seg, max_rows):
            _cabi.check(lib, handle, lib.mbexwn_gather_rows(handle, src.data_ptr(), dst.data_ptr(), row_elems, seg_dev.data_ptr(),
                                                           n_seg, max_rows, stream), "mbexwn_gather_rows")
        # the cores of a finished group travel to the host (pinned) on the auxiliary stream and are copied into the result while
        # the next group computes; only the last group's copies are exposed
        aux = engine.copy_stream()
        landed = []                                           # (event, first sample, end sample) per group

        def collect(k):
            ev, c0, c1 = landed[k]
            ev.synchronize()
            src, n_thr = bufs["out_pin"].numpy(), 4
            cuts = np.linspace(c0, c1, n_thr + 1).astype(np.int64)
            threads = [threading.Thread(target=lambda a, b: out.__setitem__(slice(a, b), src[a:b]), args=(int(cuts[i]), int(cuts[i + 1])))
                       for i in range(n_thr)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()

        for gi, grp in enumerate(groups):
            lens = [wins[i].stop - wins[i].start for i in grp]
            pb = engine.prepare_cached(lens, precision, with_noise=True, with_f0=True, with_carry=True, slot=gi & 1)
            L = pb.layout
            seg_in = np.stack([np.array([wins[i].start for i in grp], dtype=np.int64), L.utt_begin.astype(np.int64),
                               np.array(lens, dtype=np.int64)], axis=1)
            seg_out = np.stack([L.utt_begin.astype(np.int64) + np.array([wins[i].core0 - wins[i].start for i in grp], dtype=np.int64),
                                np.array([wins[i].core0 for i in grp], dtype=np.int64),
                                np.array([wins[i].core1 - wins[i].core0 for i in grp], dtype=np.int64)], axis=1)
            seg = torch.from_numpy(np.concatenate([seg_in, seg_out]).reshape(-1)).to(dev, non_blocking=True)
            n = len(grp)
            seg_i, seg_o = seg[:3 * n], seg[3 * n:]
            gather(mel_full, pb.mel_dev, plan.mel_channels, seg_i, n, max(lens))
            gather(bufs["noise"], pb.noise_dev, spf, seg_i, n, max(lens))
            gather(bufs["f0"], pb.f0_dev, ppf, seg_i, n, max(lens))
            idx = torch.from_numpy(np.array([wins[i].start // align for i in grp], dtype=np.int64)).to(dev, non_blocking=True)
            pb.carry_dev[:n].copy_(bufs["run"].index_select(0, idx))
            pb.run_device()
            gather(pb.out_dev, bufs["out"], hop, seg_o, n, max(w.core1 - w.core0 for w in (wins[i] for i in grp)))
            c0, c1 = wins[grp[0]].core0 * hop, wins[grp[-1]].core1 * hop        # the windows of a group are consecutive
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
            aux.wait_event(done)
            with torch.cuda.stream(aux):
                bufs["out_pin"][c0:c1].copy_(bufs["out"][c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(aux)
            landed.append((ev, c0, c1))
            if gi == 0 and not first_done:
                # first audio of a call without a separate first window: the first group's cores
                ev.synchronize()
                info["first_chunk_latency_s"] = time.perf_counter() - t0
            if gi > 0:
                collect(gi - 1)
        if groups:
            collect(len(groups) - 1)
        info["main_pass_s"] = time.perf_counter() - t1
    info["total_s"] = time.perf_counter() - t0
    info["audio_s"] = T * hop / plan.sample_rate
    info["n_windows"] = len(wins)
    info["context_frames"] = ctx
    return out, info
