"""Init-time DSP constants for the MBExWN forward path (CPU, NumPy/SciPy, run once at model load).

Everything here runs in float64 on the host when a model is created and is then
frozen into float32 device constants:

* the band-limited Liljencrants-Fant glottal-pulse wavetables
  (reference: MBExWN_NVoc/vocoder/model/tf_wavetable.py:37-162, :182-307, :310-410,
  MBExWN_NVoc/glottis/FglotspecLF.py:15-217, MBExWN_NVoc/glottis/FglotLFsynthparams.py:12-191)
* the PQMF prototype and the cosine-modulated synthesis/analysis banks
  (reference: MBExWN_NVoc/vocoder/model/tf_preprocess.py:30-80, :120-161)
* the STFT analysis window and its overlap-add dual
  (reference call sites: custom_pulsed_generator.py:388-400, :692-694, :716-724)
* the F0-dependent cepstral lifter bank and the F0 smoothing kernel
  (reference: custom_pulsed_generator.py:403-406, :434-450)

None of this is on the per-utterance hot path; it is the part of the model
constructor that is arithmetic rather than plumbing.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import scipy.optimize
import scipy.signal
import scipy.signal.windows


# --------------------------------------------------------------------------------------
# Liljencrants-Fant glottal pulse: synthesis parameters and analytic spectrum
# --------------------------------------------------------------------------------------

def _bracket_and_solve(fun, what: str) -> float:
    """Root of `fun` near 0: grow a bracket outwards in unit steps on both signs, then Brent.

    Mirrors the bracketing strategy of FglotLFsynthparams.py:83-103 / :168-187 so the same
    root is selected when several exist.
    """
    f0 = fun(0.0)
    lo, hi = 0.0, 0.1
    if abs(f0) > np.finfo(np.float64).eps:
        while f0 * fun(hi) > 0 and f0 * fun(-hi) > 0:
            lo = hi
            hi += 1.0
        if fun(-hi) * f0 < 0:
            lo, hi = -lo, -hi
    else:
        lo, hi = -0.1, 0.1
    root = scipy.optimize.brentq(fun, lo, hi)
    if root > max(lo, hi):
        raise RuntimeError(f"LF model: {what} did not converge")
    return float(root)


def lf_synthesis_params(oq: float, am: float, ta: float) -> Tuple[float, float, float]:
    """Solve the LF-model growth factor `alpha` and return-phase parameter `epar`.

    Follows FglotLFsynthparams.py:12-191 (non-`old` branches). Times are relative to a unit
    period: te = oq, tp = am*oq, wg = pi/tp. Returns (alpha, epar, ta) where ta may be
    adjusted for the degenerate ranges exactly as the reference does.
    """
    eps = np.finfo(np.float64).eps
    if oq <= eps or oq >= 1 - eps:
        raise RuntimeError("open quotient out of range")
    if am < 0.5 or am >= 1 - eps:
        raise RuntimeError("asymetry is out of range")
    if ta < 0 or ta > 1 - oq:
        raise RuntimeError("return phase length(ta) is out of range")

    te = oq
    wg = np.pi / (oq * am)
    c = np.cos(wg * te)
    s = np.sin(wg * te)

    if ta <= np.finfo(np.float32).eps:
        # abrupt closure: the open-phase flow derivative must integrate to zero on its own
        def eq_alpha(a):
            return np.exp(a * oq) * (wg * c - a * s) - wg
        return _bracket_and_solve(eq_alpha, "alpha"), 0.0, 0.0

    if oq > 0.999:
        epar, ta = 0.5, 0.5 * (1 - oq)
    elif ta > 0.99 * (1 - oq):
        epar, ta = 0.0, 1 - oq
    else:
        q = (te - 1) / ta

        def eq_epar(e):
            return e - 1 + np.exp(e * q)
        e_left = -np.log(-q) / q
        epar = float(scipy.optimize.brentq(eq_epar, e_left, 1.1))

    # area under the return phase (closed form), then alpha balances the open phase against it
    if epar == 0:
        ret_area = -ta / 2
    else:
        ex = np.exp(epar / ta * (te - 1))
        ret_area = (-ex * (ta + epar - te * epar) + ta) / (epar * (-1 + ex))

    def eq_alpha(a):
        return -(-wg * c + a * s + wg * np.exp(-a * te)) / (a * a + wg * wg) / s + ret_area

    return _bracket_and_solve(eq_alpha, f"alpha (ta={ta:f})"), float(epar), float(ta)


def _cis(x):
    return np.cos(x) + 1j * np.sin(x)


def lf_pulse_spectrum(f_harm: np.ndarray, oq: float, am: float, ta: float,
                      derivative: bool = False) -> np.ndarray:
    """Fourier transform of one LF glottal-flow(-derivative) period at normalised frequencies.

    `f_harm[k] = 1` is the fundamental. Restates FglotspecLF.py:15-217 for Ee=1, orig=0.
    With derivative=False the flow itself is returned (division by j*w, DC from the closed-form
    integrals), which is what the wavetable uses (`use_radiation=False`, tf_wavetable.py:196).
    """
    eps = np.finfo(np.float64).eps
    if oq <= eps or oq >= 1 - eps:
        raise RuntimeError(f"open quotient {oq:f} out of range")
    if am <= 0.5 or am >= 1 - eps:
        raise RuntimeError(f"asymetry {am:f} is out of range")
    if ta < 0 or ta > 1 - oq:
        raise RuntimeError(f"return phase length(ta) {ta:f} is out of range")

    te = float(oq)
    wg = np.pi / (oq * am)
    alpha, epar, ta = lf_synthesis_params(oq, am, ta)
    w = np.asarray(f_harm, dtype=np.float64) * 2 * np.pi

    # open phase: E0 exp(alpha t) sin(wg t) on [0, te], scaled so that E(te) = -1
    half_e0 = -0.5 / (np.exp(alpha * te) * np.sin(wg * te))
    grown = np.exp(alpha * te + np.log(half_e0))
    wg_eps = eps if (abs(alpha) < eps and np.min(np.abs(w - wg)) < eps) else 0.0
    spec = ((grown * _cis(te * (wg - w)) - half_e0) / (1j * alpha + (w - wg + wg_eps))
            - (grown * _cis(-te * (w + wg)) - half_e0) / (1j * alpha + (w + wg)))

    # return phase on [te, 1]
    if ta != 0:
        nz = np.flatnonzero(w > np.finfo(np.float64).eps)
        if epar > 0:
            ex = np.exp(epar * (te - 1) / ta)
            shift = _cis(-te * w)
            hh = np.ones(w.shape, dtype=np.complex128) * (-1j * (te - 1))
            hh[nz] = (shift[nz] - _cis(-w[nz])) / w[nz]
            ret = ((ta * (1 - ex)) * shift + (1j * epar * ex) * hh) \
                / (w * (1j * ta * (ex - 1)) + epar * (ex - 1))
        else:
            ret = ta * 0.5 * np.ones(w.shape, dtype=np.complex128)
            ret[nz] = (1j * ta * w[nz] - 1 + np.exp(-1j * w[nz] * ta)) / (ta * w[nz] ** 2)
            ret = ret * np.exp(-1j * oq * w)
        spec = spec + ret

    if derivative:
        if w[0] == 0:
            spec[0] = 0
        return spec

    if w[0] != 0:
        return spec / (1j * w)

    spec[1:] = spec[1:] / (1j * w[1:])
    e0 = -1.0 / (np.exp(alpha * oq) * np.sin(wg * oq))
    ea = np.exp(alpha * te)
    sw, cw = np.sin(wg * te), np.cos(wg * te)
    open_dc = e0 * (-2 * alpha * ea * wg * cw + alpha ** 2 * ea * sw - wg ** 2 * ea * sw
                    + wg * te * alpha ** 2 + wg ** 3 * te + 2 * alpha * wg) / (alpha ** 2 + wg ** 2) ** 2
    if ta > 0:
        k = epar / ta
        ex = np.exp(k * (te - 1))
        close_dc = -0.5 * ta ** 2 * (ex * (2 + k ** 2 + 2 * k + (k * te) ** 2 - 2 * k * te - 2 * k ** 2 * te) - 2) \
            / (epar ** 3)
    else:
        close_dc = 0.0
    spec[0] = open_dc + close_dc
    return spec


def pulse_lowpass_kaiser(pass_band_edge: float, stop_att_db: float = 70.0,
                         trans_width_normed: float = 0.1) -> np.ndarray:
    """Kaiser-windowed FIR low-pass used to band-limit a wavetable (tf_wavetable.py:37-80)."""
    if stop_att_db >= 50:
        beta = 0.1102 * (stop_att_db - 8.7)
    elif stop_att_db >= 21:
        beta = 0.5842 * (stop_att_db - 21.0) ** 0.4 + 0.07886 * (stop_att_db - 21.0)
    else:
        beta = 0.0
    width = 2 * np.pi * trans_width_normed
    cutoff = pass_band_edge - 0.5 * trans_width_normed
    while True:
        radius = int(np.ceil((stop_att_db - 8.0) / 2.285 / width / 2))
        if 2 * radius > 8000 and stop_att_db > 10:
            stop_att_db -= 6
        else:
            break
    return scipy.signal.firwin(2 * radius + 1, cutoff=[cutoff], window=("kaiser", beta),
                               pass_zero=True, fs=1.0)


def lf_pulse_table(n_wavetable: int, oq: float, am: float, rta: float, pul_bw: float,
                   transition_width: float, use_deriv: bool = False) -> np.ndarray:
    """One band-limited LF pulse period, synthesised in the frequency domain (tf_wavetable.py:93-162).

    The table length is the next power of two >= n_wavetable (min 16) while the pulse period stays
    n_wavetable samples, exactly as in the reference.
    """
    fft_size = 16
    while fft_size < n_wavetable:
        fft_size *= 2
    bins = np.arange(fft_size // 2 + 1) / fft_size
    spec = lf_pulse_spectrum(bins * n_wavetable, oq=oq, am=am, ta=rta * (1 - oq), derivative=use_deriv)
    fir = pulse_lowpass_kaiser(pul_bw, stop_att_db=70, trans_width_normed=min(pul_bw / 2.0, transition_width))
    over = 1
    while fir.shape[0] > fft_size * over:
        over *= 2
    fir_spec = np.fft.rfft(fir, fft_size * over)[::over]
    fir_spec[-1] = np.real(fir_spec[-1])
    return np.fft.irfft(spec * np.abs(fir_spec), fft_size)


@dataclass
class WaveTables:
    tables: np.ndarray        # (n_period + 1, K) float32; last row repeats the first
    n_period: int
    nominal_f0: float         # realised nominal F0 [Hz]
    f0_grid: List[float]
    min_transposition: np.float32
    max_transposition: np.float32
    grid_norm: np.float32     # 1 / log(F0GridFactor)
    sample_rate: float


def _normed_pulse(oq, target_nominal_f0, nominal_bw, sample_rate, am, rta, use_radiation,
                  bw_reduction, wt_oversampling):
    """tf_wavetable.py:310-410, non-sinusoid branch. Returns (pulse, realised nominal F0)."""
    n = int(np.ceil(wt_oversampling * sample_rate / target_nominal_f0))
    res = lf_pulse_table(n, oq=oq, am=am, rta=rta, pul_bw=nominal_bw / (bw_reduction * wt_oversampling),
                         transition_width=0.1 / wt_oversampling, use_deriv=use_radiation)
    return res, wt_oversampling * sample_rate / res.shape[0]


def build_wavetables(sample_rate: float, nominalF0: float, maxF0: Optional[float] = None,
                     Oq: float = 0.5, am: float = 0.8, rta: float = 0.05, use_radiation: bool = False,
                     F0GridFactor: float = 1.25, numF0InGrid: int = 5, wt_oversampling: int = 2,
                     nominalBandWidth: Optional[float] = None, **unsupported) -> WaveTables:
    """Wavetable bank of PulseWaveTable.__init__ (tf_wavetable.py:182-307).

    Options that change the run-time path (sinusoid tables, pulse-synchronous gains, trainable tables) are not part
    of the MBExWN inference path and are rejected; ``add_subharm_chans`` does not touch the tables (it adds sinusoid
    channels at run time, tf_wavetable.py:554-559) and is handled by the plan / pulse kernel.
    """
    for key in ("use_sinusoid", "use_sinusoid_as_fun", "use_white_pulse",
                "pulse_sync_gain_avg", "no_interp", "trainable"):
        if unsupported.get(key):
            raise NotImplementedError(f"wavetable_config.{key} is not supported on the B200 path")
    if maxF0 is None:
        # the reference evaluates maxF0/nominalF0 unconditionally (tf_wavetable.py:247)
        raise TypeError("wavetable_config.maxF0 is required")

    _, nominal = _normed_pulse(Oq, nominalF0, 0.5 / F0GridFactor, sample_rate, am, rta, use_radiation,
                               maxF0 / nominalF0, wt_oversampling)
    n_grid = int(np.ceil(np.log(maxF0 / nominal) / np.log(F0GridFactor))) if maxF0 is not None else numF0InGrid

    cols, grid = [], []
    for ir in range(n_grid + 1):
        rs = F0GridFactor ** ir if ir > 0 else 1
        wt, _ = _normed_pulse(Oq, nominal, 0.5, sample_rate, am, rta, use_radiation, rs, wt_oversampling)
        wt = wt.astype(np.float32)
        grid.append(nominal * rs)
        cols.append(np.concatenate([wt, wt[0:1]], axis=0)[:, np.newaxis])
    norm = -np.min([cols])                      # float32 scalar
    tables = np.concatenate([c / norm for c in cols], axis=1).astype(np.float32)
    return WaveTables(tables=np.ascontiguousarray(tables), n_period=int(tables.shape[0] - 1),
                      nominal_f0=float(nominal), f0_grid=grid,
                      min_transposition=np.float32(np.min(grid) / nominal),
                      max_transposition=np.float32(np.max(grid) / nominal),
                      grid_norm=np.float32(1.0 / np.log(np.float32(F0GridFactor))),
                      sample_rate=float(sample_rate))


# --------------------------------------------------------------------------------------
# PQMF
# --------------------------------------------------------------------------------------

def pqmf_prototype(taps: int, cutoff_ratio: float, beta: float) -> np.ndarray:
    """Kaiser-windowed sinc prototype, taps+1 coefficients (tf_preprocess.py:30-80, NumPy branch)."""
    assert taps % 2 == 0, "The number of taps mush be even number."
    assert 0.0 < cutoff_ratio < 1.0, "Cutoff ratio must be > 0.0 and < 1.0."
    n = np.arange(taps + 1) - 0.5 * taps
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.sin(np.pi * cutoff_ratio * n) / (np.pi * n)
    h[taps // 2] = cutoff_ratio
    return h * scipy.signal.windows.kaiser(taps + 1, beta)


def pqmf_filters(subbands: int, taps: int, cutoff_ratio: float, beta: float) -> Tuple[np.ndarray, np.ndarray]:
    """(analysis, synthesis) banks, each (subbands, taps+1) float32 (tf_preprocess.py:120-161)."""
    h = pqmf_prototype(taps, cutoff_ratio, beta)
    n = np.arange(taps + 1) - taps / 2
    ana = np.zeros((subbands, taps + 1))
    syn = np.zeros((subbands, taps + 1))
    for k in range(subbands):
        arg = (2 * k + 1) * (np.pi / (2 * subbands)) * n
        ana[k] = 2 * h * np.cos(arg + (-1) ** k * np.pi / 4)
        syn[k] = 2 * h * np.cos(arg - (-1) ** k * np.pi / 4)
    return ana.astype(np.float32), syn.astype(np.float32)


def pqmf_polyphase(syn: np.ndarray, subbands: int, taps: int) -> Tuple[np.ndarray, int, int]:
    """Polyphase form of TFPQMF.synthesis (tf_preprocess.py:208-226).

    y[m*S + p] = sum_{q=0}^{Q-1} sum_k x[m + q - back, k] * G[q, k, p] with
    G[q, k, p] = S * h_syn[k, S*q + off - p] where in range (zero otherwise).
    Returns (G (Q, S, S) float32, Q, back).
    """
    S = subbands
    half = taps // 2
    back = -(-half // S)                      # ceil(half / S)
    off = back * S - half                     # j = S*q - p - off ... see below
    # tap j pairs output n with stuffed sample n + j - half = t*S  =>  t = m + (p + j - half)/S
    # write t = m + q - back  =>  j = S*(q - back) + half - p
    q_max = (taps - half + (S - 1)) // S + back
    Q = q_max + 1
    G = np.zeros((Q, S, S), dtype=np.float64)
    for q in range(Q):
        for p in range(S):
            j = S * (q - back) + half - p
            if 0 <= j <= taps:
                G[q, :, p] = S * syn[:, j].astype(np.float64)
    del off
    return G.astype(np.float32), Q, back


# --------------------------------------------------------------------------------------
# STFT windows, lifters
# --------------------------------------------------------------------------------------

def hann_periodic(n: int) -> np.ndarray:
    """tf.signal.hann_window(n, periodic=True) in float32."""
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def inverse_stft_window(frame_length: int, frame_step: int) -> np.ndarray:
    """tf.signal.inverse_stft_window_fn(frame_step, hann)(frame_length) in float32.

    w_inv[n] = w[n] / sum_m w^2[(n mod step) + m*step]; call site custom_pulsed_generator.py:716-720.
    """
    w = hann_periodic(frame_length).astype(np.float32)
    overlaps = -(-frame_length // frame_step)
    padded = np.zeros(overlaps * frame_step, dtype=np.float32)
    padded[:frame_length] = w
    denom = np.square(padded).reshape(overlaps, frame_step).sum(axis=0, dtype=np.float32)
    denom = np.tile(denom, overlaps)[:frame_length]
    return (w / denom).astype(np.float32)


def stft_sizes(sample_rate: int, hop: int, internal_win_size_s: Optional[float], internal_fft_over: int):
    """custom_pulsed_generator.py:391-400."""
    win = int(internal_win_size_s * sample_rate) if internal_win_size_s else 4 * hop
    fft = 16
    while fft < win:
        fft *= 2
    return win, fft * (2 ** internal_fft_over)


def f0_smoothing_kernel(hop: int) -> np.ndarray:
    """Normalised Bartlett window without its zero end points (custom_pulsed_generator.py:405-406)."""
    w = np.bartlett(2 * hop + 3)[1:-1]
    return (w / np.sum(w)).astype(np.float32)


def cepstral_lifters(scale: float, sample_rate: int, fmin: float, fmax: float, n_ceps: int,
                     n_grid: int = 30) -> Tuple[np.ndarray, np.ndarray]:
    """F0-indexed half-Hamming lifter bank (custom_pulsed_generator.py:434-450).

    Returns (log10 f0 grid (n_grid,) float32, lifters (n_grid, n_ceps) float32).
    """
    grid, rows = [], []
    for f0 in np.logspace(np.log10(fmin), np.log10(fmax), n_grid):
        win_len = int(scale * 0.5 * sample_rate / f0)
        if win_len % 2 == 0:
            win_len += 1
        half = np.hamming(win_len)[win_len // 2:]
        if win_len // 2 + 1 > n_ceps:
            rows.append(half[:n_ceps])
        else:
            rows.append(np.concatenate((half, np.zeros(n_ceps - 1 - (win_len // 2)))))
        grid.append(np.log10(f0))
    return np.asarray(grid, dtype=np.float32), np.asarray(rows, dtype=np.float32)


# ---- analysis side: audio -> mel ------------------------------------------------------------------
def cosine_window(win_type: str, winlen: int) -> np.ndarray:
    """Symmetric 'hann' / 'hamming' window of the reference's window generator (sig_proc/Mwindows.py:61-67, 192-196):
    the first half is a1 + a2 cos(2 pi n / (N - 1)), n <= (N - 1) // 2, mirrored onto the second half."""
    coefs = {"hann": (0.5, -0.5), "hanning": (0.5, -0.5), "hamming": (0.54, -0.46)}
    if win_type.lower() not in coefs:
        raise RuntimeError("window::unsupported window type {0}".format(win_type))
    a1, a2 = coefs[win_type.lower()]
    win = np.zeros((winlen,))
    mid = (winlen - 1) // 2
    x = np.arange(mid + 1)
    half = a1 + a2 * np.cos(2. * np.pi * x / (winlen - 1)) + 0.0 * np.cos(4. * np.pi * x / (winlen - 1)) \
        + 0.0 * np.cos(6. * np.pi * x / (winlen - 1))
    win[:mid + 1] = half
    win[winlen - 1:winlen - 2 - mid:-1] = half
    return win


def _hz_to_mel_slaney(f):
    """Slaney (Auditory Toolbox) mel scale = librosa hz_to_mel(htk=False): linear below 1 kHz, log above."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_frequencies(n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.mel_frequencies(htk=False): n_mels frequencies equally spaced on the Slaney mel scale."""
    return _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels))


def mel_filter_bank(sr: float, n_fft: int, n_mels: int, fmin: float, fmax: Optional[float],
                    norm: bool = True, dtype=np.float32) -> np.ndarray:
    """(n_mels, n_fft // 2 + 1) triangular mel basis as the reference builds it with ``librosa.filters.mel(sr, n_fft,
    n_mels, fmin, fmax, htk=False, norm='slaney')`` (vocoder/model/preprocess.py:52-74, centered=False).

    librosa (>= 0.8.0, requirements.txt:5) is a third-party dependency that is absent here; this is its published
    algorithm: n_mels + 2 band edges equally spaced on the Slaney mel scale, weights
    max(0, min((f - f_lo) / (f_c - f_lo), (f_hi - f) / (f_hi - f_c))) on the FFT bin frequencies, each triangle scaled by
    2 / (f_hi - f_lo) (Slaney area normalisation)."""
    if fmax is None:
        fmax = float(sr) / 2
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm:
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis].astype(dtype)
    return weights


def mel_filter_csr(basis: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Band-compressed form of the mel basis for the analysis kernel: per band the first non-zero bin, the number of bins
    up to the last non-zero one, the offset of its weights in the packed array, and the packed weights."""
    lo, cnt, off, packed = [], [], [], []
    for row in basis:
        nz = np.flatnonzero(row)
        if nz.size == 0:
            lo.append(0), cnt.append(0), off.append(len(packed))
            continue
        lo.append(int(nz[0])), cnt.append(int(nz[-1] - nz[0] + 1)), off.append(len(packed))
        packed.extend(row[nz[0]:nz[-1] + 1].tolist())
    return (np.asarray(lo, np.int32), np.asarray(cnt, np.int32), np.asarray(off, np.int32),
            np.asarray(packed, np.float32))


def resample_filter(in_sr: int, out_sr: int, stop_att: float = 70, trans_width_normed: float = 0.1,
                    dtype=np.float64) -> Tuple[np.ndarray, int, int]:
    """Kaiser anti-aliasing FIR and the up / down factors of the reference's resampler (sig_proc/resample.py:31-63)."""
    import math
    from scipy import signal as ss
    in_sr, out_sr = int(in_sr), int(out_sr)
    gcd = math.gcd(in_sr, out_sr)
    up, down = out_sr // gcd, in_sr // gcd
    if stop_att >= 50:
        beta = 0.1102 * (stop_att - 8.7)
    elif stop_att >= 21:
        beta = 0.5842 * pow(stop_att - 21., 0.4) + 0.07886 * (stop_att - 21.)
    else:
        beta = 0.
    trans_width = 2 * np.pi * np.fmin(1., out_sr / in_sr) * trans_width_normed
    while True:
        radius = int(np.ceil((stop_att - 8.) / 2.285 / trans_width / 2))
        if (2 * radius > 8000) and stop_att > 10:
            stop_att -= 6
        else:
            break
    winlen = radius * 2 + 1
    fir = ss.firwin(winlen * up, cutoff=(1 - trans_width_normed) / max(up, down), window=("kaiser", beta))
    return fir.astype(dtype, copy=False), up, down
