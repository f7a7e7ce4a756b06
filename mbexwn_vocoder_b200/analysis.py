"""Analysis side on the GPU: audio -> log-mel, the inverse direction of the vocoder.

Host mirror of ``MELInverter.generate_mel_from_snd`` (mel_inverter.py:156-182) ->
``compute_mel_spectrogram_internal`` (vocoder/model/preprocess.py:417-560) -> ``calc_stft`` (sig_proc/spec/stft.py:14-96):
the STFT, magnitude, mel projection and log run in ONE hand-written kernel (``mel_analysis2048_kernel``, csrc/k_synth.cu)
behind ``mbexwn_mel_analysis`` (include/mbexwn.h).  PyTorch owns device / pinned memory only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi, dsp_init

MODE_LOG_FLOOR, MODE_LOG_OFFSET, MODE_LOG_MAX = 0, 1, 2


def frame_count(n_samples: int, hop: int) -> int:
    """calc_stft with center=True: len // hop + 1 frames (sig_proc/spec/stft.py:56)."""
    return n_samples // hop + 1


def analysis_layout(lengths: Sequence[int], hop: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """sample_begin [n], n_samples [n], frame_begin [n + 1], pair_first [n + 1] of a ragged batch laid back to back."""
    n = np.asarray(lengths, dtype=np.int64)
    if n.size and n.min() < 1:
        raise RuntimeError("calc_stft::error::cannot pad an empty signal")       # np.pad(mode='reflect') raises there too
    frames = n // hop + 1
    sample_begin = np.concatenate(([0], np.cumsum(n)[:-1])).astype(np.int64) if n.size else np.zeros(0, np.int64)
    frame_begin = np.concatenate(([0], np.cumsum(frames))).astype(np.int32)
    pair_first = np.concatenate(([0], np.cumsum((frames + 1) // 2))).astype(np.int32)
    return sample_begin, n.astype(np.int32), frame_begin, pair_first


class MelAnalyzer:
    """Batched audio -> log-mel on one GPU with the reference's preprocess_config keys."""

    def __init__(self, preprocess_config: Dict, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("mbexwn_vocoder_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _cabi.load()
        pc = preprocess_config
        self.pc = pc
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.sample_rate, self.hop, self.fft_size = int(pc["sample_rate"]), int(pc["hop_size"]), int(pc["fft_size"])
        self.win = int(pc["win_size"]) if "win_size" in pc else self.fft_size
        self.n_mel = int(pc["mel_channels"])
        if pc.get("norm_mel"):
            raise NotImplementedError("norm_mel pre-processing is outside the built path")
        basis = dsp_init.mel_filter_bank(self.sample_rate, self.fft_size, self.n_mel, pc["fmin"], pc["fmax"])
        lo, cnt, off, w = dsp_init.mel_filter_csr(basis)
        ang = -2.0 * np.pi * np.arange(self.fft_size // 2) / self.fft_size
        host = {"window": dsp_init.cosine_window("hann", self.win).astype(np.float32),
                "twiddle": np.stack((np.cos(ang), np.sin(ang)), axis=1).astype(np.float32),
                "mel_lo": lo, "mel_cnt": cnt, "mel_off": off, "mel_w": w}
        self._t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in host.items()}
        self.lin_amp_scale = pc["lin_amp_scale"] if pc.get("lin_amp_scale", 1) != 1 else 1
        self.lin_amp_off = pc["lin_amp_off"] if pc.get("lin_amp_off") is not None else 1.e-5
        self.mel_amp_scale = pc["mel_amp_scale"] if pc.get("mel_amp_scale", 1) != 1 else 1
        self.use_max_limit = bool(pc.get("use_max_limit", False))

    def config(self, do_post: bool) -> _cabi.AnalysisConfig:
        c = _cabi.AnalysisConfig()
        c.hop, c.win, c.fft_size, c.n_mel = self.hop, self.win, self.fft_size, self.n_mel
        if do_post:                                        # scale_mel_spectrogram (preprocess.py:80-108)
            c.mode = MODE_LOG_MAX if self.use_max_limit else MODE_LOG_OFFSET
        else:                                              # log(fmax(mel, eps)) (preprocess.py:543)
            c.mode = MODE_LOG_FLOOR
        c.lin_scale, c.lin_off, c.log_scale = float(self.lin_amp_scale), float(self.lin_amp_off), float(self.mel_amp_scale)
        c.floor = float(np.finfo(np.float32).eps)
        for k in ("window", "twiddle", "mel_lo", "mel_cnt", "mel_off", "mel_w"):
            setattr(c, k, self._t[k].data_ptr())
        return c

    def prepare(self, lengths: Sequence[int], with_mag: bool = False):
        """Device buffers and the C-ABI batch of a ragged batch (reusable across calls of the same lengths)."""
        sb, ns, fb, pf = analysis_layout(lengths, self.hop)
        dev = self.device
        st = {"sample_begin": torch.from_numpy(sb).to(dev), "n_samples": torch.from_numpy(ns).to(dev),
              "frame_begin": torch.from_numpy(fb).to(dev), "pair_first": torch.from_numpy(pf).to(dev),
              "audio": torch.empty(int(ns.sum()), dtype=torch.float32, device=dev),
              "mel": torch.empty(int(fb[-1]), self.n_mel, dtype=torch.float32, device=dev),
              "mag": torch.empty(int(fb[-1]), self.fft_size // 2 + 1, dtype=torch.float32, device=dev) if with_mag else None,
              "frame_begin_host": fb}
        b = _cabi.AnalysisBatch()
        b.n_utt, b.n_pairs, b.n_frames, b.n_samples_total = len(ns), int(pf[-1]), int(fb[-1]), int(ns.sum())
        for k in ("sample_begin", "n_samples", "frame_begin", "pair_first", "audio", "mel"):
            setattr(b, k, st[k].data_ptr())
        b.mag_tap = st["mag"].data_ptr() if with_mag else None
        st["batch"] = b
        return st

    def run(self, st, do_post: bool = False):
        """Launch on buffers from prepare() whose ``audio`` is already resident; returns the device mel tensor."""
        cfg = self.config(do_post)
        with torch.cuda.device(self.device):
            rc = self.lib.mbexwn_mel_analysis(C.byref(cfg), C.byref(st["batch"]),
                                              torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(self.lib, None, rc, "mbexwn_mel_analysis")
        return st["mel"]

    def run_host(self, st, audio_pinned: torch.Tensor, mel_pinned: torch.Tensor, do_post: bool = False):
        cfg = self.config(do_post)
        with torch.cuda.device(self.device):
            rc = self.lib.mbexwn_mel_analysis_host(C.byref(cfg), C.byref(st["batch"]), audio_pinned.data_ptr(),
                                                   mel_pinned.data_ptr(),
                                                   torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(self.lib, None, rc, "mbexwn_mel_analysis_host")
        return mel_pinned

    def __call__(self, sounds: Sequence[np.ndarray], do_post: bool = False, return_mag: bool = False):
        """List of 1-D float signals at the model sample rate -> list of (frames_u, n_mel) log-mel arrays."""
        sounds = [np.ascontiguousarray(np.asarray(s, dtype=np.float32).reshape(-1)) for s in sounds]
        st = self.prepare([s.size for s in sounds], with_mag=return_mag)
        st["audio"].copy_(torch.from_numpy(np.concatenate(sounds)))
        mel = self.run(st, do_post).cpu().numpy()
        fb = st["frame_begin_host"]
        mels = [mel[fb[i]:fb[i + 1]] for i in range(len(sounds))]
        if return_mag:
            mag = st["mag"].cpu().numpy()
            return mels, [mag[fb[i]:fb[i + 1]] for i in range(len(sounds))]
        return mels


def resample(x: np.ndarray, in_sr, out_sr, axis: int = -1) -> np.ndarray:
    """Host-side sample-rate conversion of the reference (sig_proc/resample.py:6-63): Kaiser FIR + polyphase."""
    from scipy import signal as ss
    fir, up, down = dsp_init.resample_filter(in_sr, out_sr, dtype=np.float32 if x.dtype == np.float32 else np.float64)
    return ss.resample_poly(x, up, down, axis=axis, window=fir)
