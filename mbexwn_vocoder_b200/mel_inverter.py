"""MELInverter: drop-in for MBExWN_NVoc.mel_inverter.MELInverter (mel_inverter.py:21-239) on one B200.

Same constructor, same ``scale_mel`` / ``synth_from_mel`` / ``srate`` surface and the same exception classes;
behind it the TensorFlow graph is replaced by the hand-written sm_100a kernels of libmbexwn_b200.so.  The
noise channel of the generator (custom_pulsed_generator.py:906) is drawn in-kernel from a counter-based
Philox stream seeded by ``seed`` (the reference uses TensorFlow's global generator).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
from scipy.interpolate import interp1d

from . import get_config_file, list_models  # noqa: F401  (re-exported like the reference module)
from . import config as cutils
from . import weights as W
from .plan import build_plan

log_to_db = 20 * np.log10(np.exp(1))            # vocoder/model/preprocess.py:78


SYNTHETIC_WEIGHTS_ENV = "MBEXWN_SYNTHETIC_WEIGHTS"


def resolve_weights(model_dir: str, plan, hparams: Dict, verbose: bool = False,
                    allow_synthetic_weights: Optional[bool] = None) -> Dict[str, np.ndarray]:
    """Weights of a model directory, in this order: ``weights.npz`` (this package's container), ``weights.tf``
    (the reference's TensorFlow checkpoint, mel_inverter.py:203-208, read without TensorFlow by tf_checkpoint.py).

    Without either the reference fails hard in ``load_weights`` (mel_inverter.py:203-210) and so does this function
    (FileNotFoundError) -- audio from random weights is noise, and the model directories that ship with this package hold
    synthetic configurations under the reference's model names.  Tests, benchmarks and parity runs opt in explicitly with
    ``allow_synthetic_weights=True`` (or the environment variable MBEXWN_SYNTHETIC_WEIGHTS=1): Keras-like random
    initialisation from the seed in the config, announced on stderr."""
    npz = os.path.join(model_dir, "weights.npz")
    tf_prefix = os.path.join(model_dir, "weights.tf")
    if os.path.exists(npz):
        if verbose:
            print(f"restore from {npz}", file=sys.stderr)
        weights = W.load(npz)
        W.check(plan, weights)
        return weights
    if os.path.exists(tf_prefix + ".index"):
        from .tf_checkpoint import import_weights
        if verbose:
            print(f"restore from {tf_prefix}", file=sys.stderr)
        return import_weights(tf_prefix, plan)
    if allow_synthetic_weights is None:
        allow_synthetic_weights = os.environ.get(SYNTHETIC_WEIGHTS_ENV, "") not in ("", "0")
    if not allow_synthetic_weights:
        raise FileNotFoundError(f"no weights.npz / weights.tf in {model_dir}: install the released models "
                                f"(tools/install_models.py) or pass allow_synthetic_weights=True / set {SYNTHETIC_WEIGHTS_ENV}=1 "
                                f"for random-initialised weights (tests and benchmarks only)")
    seed = int(hparams.get("synthetic_weights", {}).get("seed", 0))
    print(f"MELInverter::warning::no weights in {model_dir}: RANDOM-INITIALISED weights (seed {seed}); the output is not speech",
          file=sys.stderr)
    return W.init_synthetic(plan, seed=seed)


class MELInverter(object):
    def __init__(self, model_id_or_path: Union[str, None] = None, verbose: bool = False,
                 device: Union[int, str] = 0, precision: str = "f16f8", seed: int = 42,
                 allow_synthetic_weights: Optional[bool] = None, devices: Optional[Sequence[int]] = None):
        """Reference signature ``MELInverter(model_id_or_path=None, verbose=False)`` (mel_inverter.py:21-41) plus:

        precision   "f16f8" (default: the fp32-accurate tensor-core path the benchmarks measure), "bf16x3", "bf16", "fp32"
        device      CUDA device of ``synth_from_mel`` / ``synth_batch``
        devices     several CUDA devices for ``synth_many`` (one engine per GPU, LPT sharding, host gather)
        allow_synthetic_weights   see ``resolve_weights``"""
        self.model = None
        self.model_id_or_path = model_id_or_path
        self.config_file = None
        self.preprocess_config = None
        self.mel_channels = None
        self.hop_size = None
        self.fft_size = None
        self.fmin = None
        self.fmax = None
        self._srate = None
        self.win_len = None

        self.lin_amp_scale = 1
        self.lin_amp_off = 1.e-5
        self.mel_amp_scale = 1
        self.use_max_limit = False

        self.devices = [int(d) for d in devices] if devices else None
        self.device = self.devices[0] if self.devices else device
        self.allow_synthetic_weights = allow_synthetic_weights
        self._pool = None
        self.precision = precision
        self.seed = seed                        # bin/resynth_mel.py:65-67 seeds everything with 42
        self.plan = None
        self._analyzer = None
        if model_id_or_path:
            self.load_model(model_id_or_path=model_id_or_path, verbose=verbose)

    @property
    def srate(self):
        return self._srate

    # ------------------------------------------------------------------------------------------------
    def scale_mel(self, mel_config: Dict, verbose=False):
        """Bring an analysis dict (keys mell|mel, sr, hoplen, nfft, fmin, fmax, ...) to the model's log-mel scaling.

        Restates mel_inverter.py:48-148 (NumPy, host side).  The reference's verbose-only branches that reference
        undefined names are not reproduced.
        """
        hop_ratio = (mel_config['hoplen'] / mel_config['sr']) / (self.hop_size / self.srate)
        if verbose and np.abs(hop_ratio - 1) > 0.001:
            print(f"compensate change in analysis hop size. mel analysis has {mel_config['hoplen'] / mel_config['sr']}"
                  f" while the model expects {self.hop_size / self.srate}.", file=sys.stderr)
        if verbose and mel_config['sr'] != self.srate:
            print(f"    WARNING::sample rate of mel analysis is  {mel_config['sr']} model expects {self.srate}.",
                  file=sys.stderr)
        if mel_config['fmin'] != self.fmin:
            raise RuntimeError(f"mell fmin {mel_config['fmin']} does not match model fmin {self.fmin}")
        if ((mel_config['fmax'] is None) and self.fmax != mel_config['sr'] / 2) or \
                ((mel_config['fmax'] is not None) and mel_config['fmax'] != self.fmax):
            raise RuntimeError(f"mell fmax {mel_config['fmax']} does not match model fmax {self.fmax}")

        if "mell" in mel_config:
            log_mel = np.array(mel_config['mell'].T[np.newaxis], dtype=np.float64 if
                               mel_config['mell'].dtype == np.float64 else np.float32)
            if "log_spec_offset" in mel_config and mel_config["log_spec_offset"] != 0:
                log_mel -= mel_config["log_spec_offset"]
            if "log_spec_scale" in mel_config and mel_config["log_spec_scale"] != 1:
                log_mel /= mel_config["log_spec_scale"]
            mel = np.exp(log_mel)
        elif "mel" in mel_config:
            mel = np.array(mel_config['mel'].T[np.newaxis])
        else:
            raise RuntimeError("error::no supported mel spectrum (keys:mell or mell) in mel_config")

        n_fft = mel_config.get("nfft", None)
        if n_fft is None:
            n_fft = mel_config.get("n_fft", None)
        if n_fft is None:
            n_fft = mel_config.get("fft_size", None)
        fft_scale = self.fft_size // n_fft
        if fft_scale != 1:
            mel *= fft_scale
        if mel_config.get("lin_spec_offset") is not None and mel_config.get("lin_spec_offset", 0) != 0:
            mel -= mel_config["lin_spec_offset"]
        if "lin_spec_scale" in mel_config and mel_config["lin_spec_scale"] != 1:
            mel /= mel_config["lin_spec_scale"]
        if self.lin_amp_scale != 1:
            mel *= self.lin_amp_scale
        if self.use_max_limit:
            mell = np.log(np.fmax(mel, self.lin_amp_off)).astype(np.float32)
        else:
            mell = np.log(mel + self.lin_amp_off).astype(np.float32)

        if np.abs(hop_ratio - 1) > 0.001:
            src_hop = mel_config['hoplen'] / mel_config['sr']
            mell = interp1d(np.arange(mell.shape[1]) * src_hop, mell, axis=1, bounds_error=False,
                            fill_value="extrapolate")(
                np.arange(0, (mell.shape[1] - 1 + 0.1) * src_hop, self.hop_size / self.srate)).astype(np.float32)
        return mell * self.mel_amp_scale

    # ------------------------------------------------------------------------------------------------
    def synth_from_mel(self, scaled_mell, noise=None, seed: Optional[int] = None):
        """(B, T, n_mel) float32 log-mel -> flat float32 waveform of length B*T*hop (mel_inverter.py:151-154).

        Like the reference the batch is flattened by ``ravel``.  ``noise`` optionally supplies
        the (B, T*steps, 1) standard-normal draw of the generator's noise channel (parity runs).
        """
        scaled_mell = np.asarray(scaled_mell, dtype=np.float32)
        if scaled_mell.ndim != 3 or scaled_mell.shape[2] != self.mel_channels:
            raise RuntimeError(f"expected a (batch, frames, {self.mel_channels}) mel spectrogram, got {scaled_mell.shape}")
        noise_list = None if noise is None else [np.asarray(z) for z in noise]
        out, _ = self._forward(list(scaled_mell), noise=noise_list, seed=self.seed if seed is None else seed)
        return np.stack(out).ravel()

    def _forward(self, mels, noise=None, f0=None, seed=0, taps=()):
        """Engine.forward with a range guard for the f16f8 path: its main operand plane is fp16, so a model whose residual
        stream leaves the fp16 range (|x| > 65504; never seen with the synthetic weights) would produce non-finite samples.
        A strided probe of the output detects that and the batch is re-run on the equally accurate bf16x3 path (bf16 planes
        have the fp32 exponent range), with a note on stderr -- never a silent change of accuracy class."""
        out, tp = self.model.forward(mels, noise=noise, f0=f0, precision=self.precision, seed=seed, taps=taps)
        if self.precision != "f16f8":
            return out, tp
        # (1) the kernels that write the residual stream flag values beyond the e4m3 hi8 plane's range (|x| > 448: the
        #     correction products would silently lose their meaning, no non-finite value marks it) -- mbexwn_range_status;
        # (2) a strided probe of the output for non-finite samples (one hop in four: every frame overlaps four hops)
        flags = self.model.range_status(reset=True)
        bad = flags != 0 or not all(np.isfinite(w[::max(1, min(self.hop_size * 4 - 1, w.size // 64))]).all() for w in out)
        if bad:
            print(f"MELInverter::warning::operand range of the f16f8 path exceeded (range flags {flags}); "
                  "re-running this batch with precision bf16x3", file=sys.stderr)
            out, tp = self.model.forward(mels, noise=noise, f0=f0, precision="bf16x3", seed=seed, taps=taps)
        return out, tp

    def synth_batch(self, mels: Sequence[np.ndarray], noise=None, f0=None, taps: Sequence[str] = (),
                    seed: Optional[int] = None):
        """Variable-length batch: list of (T_u, n_mel) -> list of (T_u*hop,) waveforms [+ stage taps]."""
        out, tp = self._forward([np.asarray(m, dtype=np.float32) for m in mels], noise=noise, f0=f0,
                                seed=self.seed if seed is None else seed, taps=taps)
        return (out, tp) if taps else out

    def synth_many(self, mels: Sequence[np.ndarray], max_batch_frames: int = 32768, seed: Optional[int] = None,
                   return_stats: bool = False):
        """Any number of (T_u, n_mel) mels -> list of (T_u * hop,) waveforms in input order: the caller loop of the reference
        (bin/resynth_mel.py:72-104) as one call.  The set is LPT-sharded over ``devices`` (one engine and one host thread per
        GPU, no collective), cut into batches of at most ``max_batch_frames`` padded frames and pipelined over two pinned buffer
        sets per GPU; the result is gathered on the host (multi_gpu.py).  A waveform does not depend on the sharding."""
        from .multi_gpu import DevicePool
        if self._pool is None:
            invs = [self]
            for d in (self.devices or [self.device])[1:]:
                other = MELInverter.__new__(MELInverter)
                other.__dict__.update(self.__dict__)
                from .engine import Engine
                other.device = d
                other.model = Engine(self.plan, self.weights, device=d)
                other._pool = None
                invs.append(other)
            self._pool = DevicePool(invs)
        return self._pool.synth_many(mels, max_batch_frames, self.seed if seed is None else seed, self.precision, return_stats)

    def synth_stream(self, batches, noise=None, seed: Optional[int] = None):
        """Throughput serving: iterate over batches (each a list of (T_u, n_mel) mels) and yield, per batch, the list of
        waveforms.  Consecutive batches alternate between two device buffer sets, so the host->device copy of the next
        batch and the device->host copy of the previous one run under the kernels of the current one
        (mbexwn_forward_host_begin / _wait).  ``noise``: optional iterable of per-batch noise lists (parity runs)."""
        eng = self.model
        seed = self.seed if seed is None else seed
        pending = []                                         # (slot, PreparedBatch, mels, noise, precision)
        noise_it = iter(noise) if noise is not None else None
        # f16f8 range guard (mbexwn_range_status): the word is sticky and shared by the two batches in flight, so once it is
        # raised every f16f8 batch still in flight is re-run on bf16x3 and the rest of the stream is submitted as bf16x3
        state = {"prec": self.precision, "warned": False}
        for i, mels in enumerate(batches):
            slot = i & 1
            if len(pending) == 2:                            # the slot about to be reused must be drained first
                yield self._drain_stream(pending.pop(0), seed, state)
            nz = next(noise_it) if noise_it is not None else None
            mels = [np.asarray(m, dtype=np.float32) for m in mels]
            pb = eng.prepare_cached([m.shape[0] for m in mels], state["prec"], nz is not None, slot=slot)
            pb.load(mels, nz)
            pb.begin_host(slot, seed)
            pending.append((slot, pb, mels, nz, state["prec"]))
        for item in pending:
            yield self._drain_stream(item, seed, state)
        if self.precision == "f16f8":
            eng.range_status(reset=True)

    def _drain_stream(self, item, seed, state):
        slot, pb, mels, nz, prec = item
        pb.wait_host(slot)
        if prec == "f16f8" and self.model.range_status(reset=False) != 0:
            if not state["warned"]:
                print("MELInverter::warning::operand range of the f16f8 path exceeded; this stream continues with precision bf16x3",
                      file=sys.stderr)
                state["warned"] = True
            state["prec"] = "bf16x3"
            return self.model.forward(mels, noise=nz, precision="bf16x3", seed=seed)[0]
        return pb.waveforms_copy()

    def synth_long_from_mel(self, scaled_mell, noise=None, chunk_frames: int = 400, max_batch_frames: int = 32768,
                            seed: Optional[int] = None, return_info: bool = False):
        """One long (T, n_mel) or (1, T, n_mel) mel -> flat waveform, synthesised in windows of `chunk_frames` frames with
        receptive-field overlap and a carried pulse phase (long_form.py), device memory bounded by `max_batch_frames` instead
        of T.  With the same explicit `noise` (T * steps values) it equals synth_from_mel / synth_batch on the same input bit
        for bit; with noise=None the draw comes from NumPy's default_rng(seed) on the host, whereas synth_from_mel uses the
        in-kernel Philox stream -- the two then differ in their noise channel (same distribution, different numbers)."""
        from .long_form import synth_long
        mel = np.asarray(scaled_mell, dtype=np.float32)
        if mel.ndim == 3:
            if mel.shape[0] != 1:
                raise RuntimeError("synth_long_from_mel takes one utterance")
            mel = mel[0]
        if mel.ndim != 2 or mel.shape[1] != self.mel_channels:
            raise RuntimeError(f"expected a (frames, {self.mel_channels}) mel spectrogram, got {mel.shape}")
        if noise is None:
            rng = np.random.default_rng(self.seed if seed is None else seed)
            noise = rng.standard_normal(mel.shape[0] * self.plan.steps_per_frame, dtype=np.float32)
        out, info = synth_long(self.model, mel, np.asarray(noise).reshape(-1), chunk_frames, self.precision, max_batch_frames)
        return (out, info) if return_info else out

    def generate_mel_from_snd(self, snd, srate):
        """Audio -> analysis dict with the log-mel under ``mell`` (mel_inverter.py:156-182); feed it to ``scale_mel``.

        The STFT / mel projection / log run on the GPU (analysis.MelAnalyzer); resampling to the model rate, when
        needed, stays on the host like in the reference (sig_proc/resample.py)."""
        from .analysis import MelAnalyzer, resample
        data_dict = {'nfft': self.fft_size, 'hoplen': self.hop_size, 'winlen': self.win_len, 'nmels': self.mel_channels,
                     'sr': self.srate, 'fmin': self.fmin, 'fmax': self.fmax, 'lin_spec_offset': self.lin_amp_off,
                     'lin_spec_scale': self.lin_amp_scale, 'log_spec_offset': 0., 'log_spec_scale': self.mel_amp_scale,
                     "time_axis": 1}
        snd = np.asarray(snd)
        if srate != self.srate:
            snd = resample(snd, srate, self.srate, axis=-1)
        if len(snd.shape) == 1:
            snd = np.array(snd)[np.newaxis]
        if self._analyzer is None:
            self._analyzer = MelAnalyzer(self.preprocess_config, device=self.device)
        mel_ref = self._analyzer([snd[0]], do_post=False)[0]
        data_dict['mell'] = mel_ref.T
        return data_dict

    # ------------------------------------------------------------------------------------------------
    def load_model(self, model_id_or_path, verbose=False):
        """Config lookup, plan, weights and engine construction (mel_inverter.py:184-239)."""
        from .engine import Engine

        config_file = get_config_file(model_id_or_path=model_id_or_path)
        model_dir = os.path.dirname(config_file)
        hparams = cutils.read_config(config_file=config_file)
        self.config_file = config_file
        self.preprocess_config = hparams["preprocess_config"]
        self.plan = build_plan(hparams)

        weights = resolve_weights(model_dir, self.plan, hparams, verbose=verbose, allow_synthetic_weights=self.allow_synthetic_weights)
        self.weights = weights
        self.model = Engine(self.plan, weights, device=self.device)

        pc = self.preprocess_config
        self.mel_channels = pc["mel_channels"]
        self.hop_size = pc["hop_size"]
        self.fft_size = pc["fft_size"]
        self.fmin = pc["fmin"]
        self.fmax = pc["fmax"]
        self._srate = pc['sample_rate']
        self.win_len = pc['win_size'] if 'win_size' in pc else self.fft_size
        self.lin_amp_scale = pc["lin_amp_scale"] if pc.get("lin_amp_scale", 1) != 1 else 1
        self.lin_amp_off = pc["lin_amp_off"] if pc.get("lin_amp_off") is not None else 1.e-5
        self.mel_amp_scale = pc["mel_amp_scale"] if pc.get("mel_amp_scale", 1) != 1 else 1
        self.use_max_limit = bool(pc.get("use_max_limit", False))
        return
