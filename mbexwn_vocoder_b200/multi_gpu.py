"""Many utterances on one or several B200s: the caller loop of the reference (bin/resynth_mel.py:72-104 -- one
``synth_from_mel`` call per file) as a sharded, batched, pipelined run.

Utterances are independent (custom_pulsed_generator.py:556-771 has no cross-utterance op), so the set is split over the
GPUs by longest-processing-time-first bin packing of the frame counts (``sched.lpt_shards``) with **no collective on the
data path**; what comes back is a host gather: every waveform lands in the caller's (pinned) host memory and the result is
ONE list in input order.  On a GPU the shard is cut into batches of at most ``max_batch_frames`` padded frames (LPT order,
so a batch holds utterances of similar length), and consecutive batches alternate between two buffer sets of fixed
capacity: the host scatters batch i + 1 into its pinned grid and gathers batch i - 1 out of its pinned grid while the GPU
computes batch i (``mbexwn_forward_host_begin`` / ``_wait``; the copies ride on two side streams).

The in-kernel noise stream is keyed by (seed, global utterance id, position), so a waveform does not depend on the number
of shards, on the batch it travelled in or on the GPU that computed it (tests/test_gpu_multi.py: bit-identical).

Two front ends:
  * ``DevicePool`` -- one process, one engine and one host thread per GPU (MELInverter(devices=[...]).synth_many);
  * ``run_shard``  -- one rank of a torchrun job (bench.py --workload config4 --gpus N): rank r computes shard r of the same
                      deterministic LPT split; the ranks' result tables are gathered on rank 0.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import sched


def plan_batches(lengths: Sequence[int], ids: Sequence[int], halo: int, max_batch_frames: int, max_batch_utts: int = 4096) -> List[List[int]]:
    """Cut the utterances `ids` (frame counts `lengths[i]`) into batches of at most `max_batch_frames` padded frames, longest first."""
    order = sorted(ids, key=lambda i: (-int(lengths[i]), i))
    batches, cur, frames = [], [], halo
    for i in order:
        n = int(lengths[i]) + halo
        if cur and (frames + n > max_batch_frames or len(cur) >= max_batch_utts):
            batches.append(cur)
            cur, frames = [], halo
        cur.append(i)
        frames += n
    if cur:
        batches.append(cur)
    return batches


@dataclass
class ShardStats:
    device: int = 0
    n_utts: int = 0
    frames: int = 0
    batches: int = 0
    wall_s: float = 0.0
    host_prep_s: float = 0.0        # scatter of the mels into the pinned grids
    host_gather_s: float = 0.0      # copies of the waveforms out of the pinned grids
    wait_s: float = 0.0             # host blocked on the GPU
    range_reruns: int = 0
    extra: Dict[str, float] = field(default_factory=dict)


def _copy_all(pool, jobs):
    """jobs: list of (dst_view, src) -- NumPy copies release the GIL, so a few host threads move the utterances in parallel."""
    if pool is None or len(jobs) < 8:
        for dst, src in jobs:
            dst[...] = src
        return
    n = pool._max_workers

    def part(k):
        for dst, src in jobs[k::n]:
            dst[...] = src
    list(pool.map(part, range(n)))


def shard_capacity(lengths: Sequence[int], ids: Sequence[int], halo: int, max_batch_frames: int):
    """(padded frames, utterances) of the largest batch plan_batches cuts out of `ids`: the capacity of the two buffer sets."""
    batches = plan_batches(lengths, list(ids), halo, max_batch_frames)
    return (max(sum(int(lengths[i]) + halo for i in b) + halo for b in batches), max(len(b) for b in batches))


def run_shard(inv, get_mel: Callable[[int], np.ndarray], lengths: Sequence[int], ids: Sequence[int], out,
              max_batch_frames: int = 32768, seed: int = 0, precision: Optional[str] = None, keep: bool = True,
              host_threads: int = 4, capacity=None) -> ShardStats:
    """Synthesize the utterances `ids` on `inv`'s GPU: pipelined batches, waveforms gathered on the host.

    get_mel(i) returns the (T_i, n_mel) float32 mel of utterance i (a view is fine: it is copied into the pinned grid).
    out: a dict (out[id] = a fresh copy of the waveform) or a callable out(id, view) that returns the destination array the
    view of the pinned grid is to be copied into (a caller that owns one big result buffer returns its slice), or None after
    consuming the view itself.
    keep = False drops the waveforms after touching them (warm-up).
    host_threads: threads that scatter the mels into / gather the waveforms out of the pinned grids (memcpy bound).
    capacity: (padded frames, utterances) to size the two buffer sets for, if larger than this call needs (shard_capacity)."""
    import torch
    eng, plan = inv.model, inv.plan
    precision = precision or inv.precision
    st = ShardStats(device=int(eng.device.index if hasattr(eng.device, "index") and eng.device.index is not None else 0))
    ids = list(ids)
    if not ids:
        return st
    t_start = time.perf_counter()
    batches = plan_batches(lengths, ids, eng.halo, max_batch_frames)
    cap_frames = max(sum(int(lengths[i]) + eng.halo for i in b) + eng.halo for b in batches)
    cap_utts = max(len(b) for b in batches)
    if capacity is not None:                        # buffer sets sized for a larger set (a warm-up that allocates for the real run)
        cap_frames, cap_utts = max(cap_frames, int(capacity[0])), max(cap_utts, int(capacity[1]))
    key = ("multi_gpu", precision, cap_frames, cap_utts)
    slots = getattr(eng, "_shard_slots", {}).get(key)
    if slots is None:
        with torch.cuda.device(eng.device):
            slots = [eng.prepare([1], precision=precision, with_noise=False, capacity_frames=cap_frames, capacity_utts=cap_utts)
                     for _ in range(2)]
        eng._shard_slots = {key: slots}             # one geometry is kept: a serving loop calls with the same budget
    in_flight: List[Optional[List[int]]] = [None, None]
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=host_threads) if host_threads > 1 else None
    hop = plan.hop

    def drain(s: int):
        grp = in_flight[s]
        if grp is None:
            return
        t0 = time.perf_counter()
        slots[s].wait_host(s)
        t1 = time.perf_counter()
        st.wait_s += t1 - t0
        # the range-guard word is sticky and shared by the two batches in flight: once raised, every f16f8 batch that is
        # drained is re-run on bf16x3 (same accuracy class, fp32 exponent range); the word is cleared when the shard is done
        flagged = precision == "f16f8" and eng.range_status(reset=False) != 0
        if flagged:
            st.range_reruns += 1
            mels = [get_mel(i) for i in grp]
            waves, _ = eng.forward(mels, precision="bf16x3", seed=seed, utt_ids=grp)
        else:
            waves = slots[s].waveforms()
        if keep and callable(out):
            jobs = [out(i, w) for i, w in zip(grp, waves)]        # a callable returns the destination view (or copies itself)
            jobs = [(d, w) for d, w in zip(jobs, waves) if d is not None]
            _copy_all(pool, jobs)
        elif keep:
            dsts = [np.empty(int(lengths[i]) * hop, dtype=np.float32) for i in grp]
            _copy_all(pool, list(zip(dsts, waves)))
            for i, d in zip(grp, dsts):
                out[i] = d
        else:
            for i, w in zip(grp, waves):
                _ = float(w[0])                     # touch only
        st.host_gather_s += time.perf_counter() - t1
        in_flight[s] = None

    with torch.cuda.device(eng.device):
        for k, grp in enumerate(batches):
            s = k & 1
            drain(s)
            t0 = time.perf_counter()
            pb = slots[s].rebind([int(lengths[i]) for i in grp])
            pb.set_utt_ids(grp)
            L, mh = pb.layout, pb.mel_host.numpy()
            _copy_all(pool, [(mh[L.utt_begin[k]:L.utt_end[k]], get_mel(i)) for k, i in enumerate(grp)])
            st.host_prep_s += time.perf_counter() - t0
            pb.begin_host(s, seed=seed)
            in_flight[s] = grp
            st.frames += int(sum(int(lengths[i]) for i in grp))
        # drain in submission order
        last = (len(batches) - 1) & 1
        drain(1 - last)
        drain(last)
        if precision == "f16f8":
            eng.range_status(reset=True)
    if pool is not None:
        pool.shutdown()
    st.n_utts, st.batches = len(ids), len(batches)
    st.wall_s = time.perf_counter() - t_start
    return st


class DevicePool:
    """One MELInverter-like engine per GPU inside one process, one host thread each."""

    def __init__(self, inverters: Sequence):
        self.inverters = list(inverters)

    def synth_many(self, mels: Sequence[np.ndarray], max_batch_frames: int = 32768, seed: int = 0,
                   precision: Optional[str] = None, return_stats: bool = False):
        lengths = [int(np.asarray(m).shape[0]) for m in mels]
        n = len(self.inverters)
        shards = sched.lpt_shards(lengths, n)
        import os
        host_threads = max(1, min(4, (os.cpu_count() or 4) // n))     # copy threads per GPU: the host cores are shared
        out: Dict[int, np.ndarray] = {}
        stats: List[Optional[ShardStats]] = [None] * n
        errors: List[BaseException] = []

        def work(k: int):
            try:
                stats[k] = run_shard(self.inverters[k], lambda i: np.asarray(mels[i], dtype=np.float32), lengths, shards[k], out,
                                     max_batch_frames, seed, precision, host_threads=host_threads)
            except BaseException as e:          # surfaced in the caller's thread
                errors.append(e)

        if n == 1:
            work(0)
        else:
            threads = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(n)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        result = [out[i] for i in range(len(mels))]             # the host gather: one list, input order
        return (result, stats) if return_stats else result


def imbalance(lengths: Sequence[int], shards: Sequence[Sequence[int]]) -> float:
    """max shard load / mean shard load of an LPT split (1.0 = perfectly even)."""
    loads = [sum(int(lengths[i]) for i in s) for s in shards]
    return max(loads) / (sum(loads) / len(loads)) if sum(loads) else 1.0
