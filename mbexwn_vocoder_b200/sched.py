"""Batch geometry: the padded frame grid, and utterance -> GPU sharding.

The reference processes one utterance per call with padding at that utterance's own ends.
To batch utterances of unequal length without changing a single output sample, they are laid out back to back
on one time axis with `halo` all-zero guard frames around each; every kernel resolves its boundary rule from
the per-frame utterance id.  Multi-GPU: utterances are independent (custom_pulsed_generator.py:556-771 has no
cross-utterance op), so a batch is split by longest-processing-time bin packing over frame counts with no
collective on the data path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


@dataclass
class FrameGridLayout:
    n_utt: int
    n_frames: int               # padded frames including guards
    n_chunks: int
    halo: int
    lengths: np.ndarray         # (n_utt,) frames per utterance
    frame_utt: np.ndarray       # (n_frames,) int32, -1 on guard frames
    utt_begin: np.ndarray       # (n_utt,) int32
    utt_end: np.ndarray         # (n_utt,) int32
    chunk_first: np.ndarray     # (n_utt + 1,) int32

    def scatter(self, per_utt: Sequence[np.ndarray], rate: int, out: np.ndarray) -> np.ndarray:
        """Copy per-utterance arrays (T_u * rate, ...) into a zeroed grid buffer (n_frames * rate, ...)."""
        for u, a in enumerate(per_utt):
            out[self.utt_begin[u] * rate:self.utt_end[u] * rate] = a
        return out

    def gather(self, grid_buf: np.ndarray, rate: int) -> List[np.ndarray]:
        return [grid_buf[self.utt_begin[u] * rate:self.utt_end[u] * rate] for u in range(self.n_utt)]


def make_layout(lengths: Sequence[int], halo: int, pulse_per_frame: int, chunk: int = 1000) -> FrameGridLayout:
    lengths = np.asarray(lengths, dtype=np.int64)
    if lengths.ndim != 1 or lengths.size == 0 or np.any(lengths <= 0):
        raise ValueError("every utterance needs at least one mel frame")
    n = lengths.size
    begin = halo + np.concatenate(([0], np.cumsum(lengths[:-1] + halo)))
    end = begin + lengths
    n_frames = int(end[-1] + halo)
    frame_utt = np.full(n_frames, -1, dtype=np.int32)
    for u in range(n):
        frame_utt[begin[u]:end[u]] = u
    chunks = -(-(lengths * pulse_per_frame) // chunk)
    chunk_first = np.concatenate(([0], np.cumsum(chunks))).astype(np.int32)
    return FrameGridLayout(n_utt=n, n_frames=n_frames, n_chunks=int(chunk_first[-1]), halo=halo, lengths=lengths,
                           frame_utt=frame_utt, utt_begin=begin.astype(np.int32), utt_end=end.astype(np.int32),
                           chunk_first=chunk_first)


def lpt_shards(lengths: Sequence[int], n_shards: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterance indices to shards (cost = frame count).

    Deterministic: ties broken by utterance index, shards by lowest load then lowest id.
    """
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * n_shards
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for i in order:
        k = min(range(n_shards), key=lambda s: (loads[s], s))
        shards[k].append(i)
        loads[k] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards
