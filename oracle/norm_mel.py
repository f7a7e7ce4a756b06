"""CPU ORACLE of NormMelComponents (wavegen_1d.py:578-769)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The optional mel-derived RMS normaliser of PaNWaveNet.infer (wavegen_1d.py:493-512): the signal RMS
is estimated per frame from the log-mel, smoothed by overlap-adding (squared) Hann windows at the sample rate, the mel is
divided by the smoothed frame RMS before the generator and the generated signal is multiplied by the sample-rate RMS
afterwards.  NumPy float32, op by op like the TensorFlow code.  Unpinned against a TensorFlow run (TensorFlow absent); pinned
on the reference's own source executed over NumPy stand-ins for the TensorFlow primitives
(tests/golden/make_reference_norm_goldens.py -> tests/test_reference_source.py) and anchored on closed forms in tests/test_norm_mel.py.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from mbexwn_vocoder_b200 import dsp_init

KERAS_EPSILON = np.float32(1e-7)                 # tf.keras.backend.epsilon()


def overlap_and_add(frames: np.ndarray, hop: int) -> np.ndarray:
    """tf.signal.overlap_and_add: (B, F, W) -> (B, (F - 1) hop + W)."""
    B, F, W = frames.shape
    out = np.zeros((B, (F - 1) * hop + W), dtype=frames.dtype)
    for f in range(F):
        out[:, f * hop:f * hop + W] += frames[:, f]
    return out


class OracleNormMel:
    def __init__(self, preprocess_config: Dict, model_config: Dict):
        pc, mc = preprocess_config, model_config
        self.win = pc.get("win_size", pc["fft_size"])
        self.hop = pc["hop_size"]
        if 4 * self.hop != self.win:                                                       # :592-594
            raise RuntimeError("NormMelComponents:error: this module currently supports only the case where "
                               "win_size {win_size} = 4 * hop_size {hop_size}")
        self.rms_norm_fact = pc["fft_size"] * self.win * 0.5                               # :598
        self.n_mel = pc["mel_channels"]
        self.use_pinv = bool(mc.get("normalize_use_pinv", False))
        hann = dsp_init.cosine_window("hann", self.win).astype(np.float32)
        if self.use_pinv:                                                                  # :602-608
            self.win_norm = np.sqrt(np.sum(hann ** 2))
            basis = dsp_init.mel_filter_bank(pc["sample_rate"], pc["fft_size"], self.n_mel, pc["fmin"], pc["fmax"])
            self.mel_band_filter_inverted = np.linalg.pinv(basis).T
        else:                                                                              # :609-612
            mel_f = dsp_init.mel_frequencies(self.n_mel + 2, pc["fmin"], pc["fmax"])
            self.inv_enorm = ((mel_f[2:self.n_mel + 2] - mel_f[:self.n_mel]) / 2.).astype(np.float32)
        self.lin_amp_scale = mc.get("lin_amp_scale", 1.)
        self.lin_amp_off = mc.get("lin_amp_off", 1.e-5)
        self.mel_amp_scale = mc.get("mel_amp_scale", 1.)
        self.use_max_limit = mc.get("use_max_limit", False)
        self.max_norm_fact = mc.get("max_norm_fact")
        self.compressor_exp = mc.get("normalize_compressor_exp")
        self.gwin = (hann / np.sum(hann))                                                  # :622-623
        scale = mc.get("normalize_smooth_win_scale", 1)
        self.smooth_win_size = int(self.win * scale)
        self.smooth_syn_win = dsp_init.cosine_window("hann", self.smooth_win_size).astype(np.float32)
        if mc.get("normalize_smooth_with_squared_win", True):
            self.smooth_syn_win = self.smooth_syn_win ** 2
        self.iters = int(mc.get("normalize_rms_num_smooth_iters", 0))
        if self.iters <= 0:
            raise NotImplementedError("normalize_rms_num_smooth_iters = 0 (per-channel time average, wavegen_1d.py:722) "
                                      "is outside the restated path")

    def frame_rms(self, mell: np.ndarray) -> np.ndarray:
        """(B, T, n_mel) log-mel -> (B, T) raw frame RMS estimate (wavegen_1d.py:663-690)."""
        mel = np.exp(mell.astype(np.float32))
        if self.use_pinv:
            test = np.tensordot(mel, self.mel_band_filter_inverted, axes=1) / self.win_norm
            rms = np.sqrt(np.sum(np.square(test), axis=-1) / self.rms_norm_fact)
        else:
            rms = np.sqrt(np.sum(np.square(mel * self.inv_enorm), axis=-1) / self.rms_norm_fact)
        rms = rms.astype(np.float32)
        if self.max_norm_fact:
            rms = np.maximum(rms, np.float32(1. / self.max_norm_fact))
        if self.compressor_exp is not None:
            rms = np.power(rms, np.float32(self.compressor_exp))
        return rms

    def normalize_inputs_by_rms(self, mell: np.ndarray, synth_length: int) -> Tuple[np.ndarray, np.ndarray, Dict]:
        """-> (normalised log-mel (B, T, n_mel), upsampled_rms (B, synth_length), taps); wavegen_1d.py:638-769, audio=None."""
        mell = np.asarray(mell, dtype=np.float32)
        B, T, _ = mell.shape
        mel = np.exp(mell)
        rms = self.frame_rms(mell)
        taps = {"rms_raw": rms.copy()}
        off = self.smooth_win_size // 2 + 2 * self.hop - self.win // 2                    # :701, :712
        ones = np.ones((1, T + 4), dtype=np.float32)
        norm_gain = overlap_and_add(ones[:, :, None] * self.smooth_syn_win[None, None, :], self.hop)[:, off:]
        gain = None
        for _ in range(self.iters):
            padded = np.concatenate((rms[:, :1], rms[:, :1], rms, rms[:, -1:], rms[:, -1:]), axis=1)
            gain = overlap_and_add(padded[:, :, None] * self.smooth_syn_win[None, None, :], self.hop)[:, off:]
            gain = gain / np.maximum(KERAS_EPSILON, norm_gain)
            n_out = (gain.shape[1] - self.win) // self.hop + 1                             # conv1d VALID, stride hop
            idx = np.arange(n_out)[:, None] * self.hop + np.arange(self.win)[None, :]
            rms = (gain[:, idx] * self.gwin[None, None, :]).sum(axis=-1, dtype=np.float32)[:, :T]
        taps["rms"] = rms.copy()
        mel = mel / np.maximum(KERAS_EPSILON, rms[:, :, None]) * np.float32(self.lin_amp_scale)
        if self.use_max_limit:
            out_mell = np.float32(self.mel_amp_scale) * np.log(np.maximum(mel, np.float32(self.lin_amp_off)))
        else:
            out_mell = np.float32(self.mel_amp_scale) * np.log(mel + np.float32(self.lin_amp_off))
        gain_off = self.win // 2
        up = np.maximum(gain[:, gain_off:gain_off + synth_length], KERAS_EPSILON)
        if up.shape[1] < synth_length:                                                     # :762-766
            up = np.concatenate((up, np.repeat(up[:, -1:], synth_length - up.shape[1], axis=1)), axis=1)
        return out_mell.astype(np.float32), up.astype(np.float32), taps
