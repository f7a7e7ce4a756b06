"""CPU ORACLE of the MBExWN mel-inversion forward pass  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product (``mbexwn_vocoder_b200``) never does; it fails loudly
when its CUDA library is missing.

What this is
------------
A NumPy / torch-CPU restatement of the reference's TensorFlow forward, stage by stage, with every
intermediate exposed as a tap.  Each function cites the reference file:line it follows
(paths relative to /root/reference/MBExWN_NVoc/).  dtype is switchable (float32 = the reference's
arithmetic, float64 = a noise-floor estimate for the tolerances used in the parity tests).

Parity pinning status (read this before trusting it)
-----------------------------------------------------
The reference ships no tests, no golden vectors, no config and no weights, and its runtime (TensorFlow)
is not installable here, so the restatement below is **unpinned against a TensorFlow run** of the reference.
What *is* pinned against real reference code executed in the build container:

* init-time DSP (tests/golden/make_reference_goldens.py -> reference_init_dsp.npz): the LF glottal-pulse
  model, the wavetable bank and the PQMF prototype, bit for bit.  Those constants come from
  ``mbexwn_vocoder_b200.dsp_init`` (the oracle imports the product for them, never the other way round).
* the forward's *algorithm as the reference wrote it*: tests/golden/make_reference_{pulse,excitation,forward,model,norm}_goldens.py compile
  the reference's own source unmodified -- ``MBExWN.call`` (inference branch), ``generate_subnet_from_specs``, ``generate_f0``,
  ``generate_excitation``, ``generate_specenv``, ``_get_cepstral_windows``, ``PulseWaveTable.call`` / ``stable_cumsum_and_wrap`` /
  ``_linear_lookup``, ``WaveNetAE(.Block).call``, the ``call`` methods of the weight-norm / sub-pixel conv, LinInterp, pad and
  activation layers, ``TFPQMF`` -- and execute it over NumPy float32 stand-ins for the TensorFlow *primitives* (conv1d, cumsum,
  gather, pad, rfft, tf.signal.stft / inverse_stft from their documentation ...).  tests/test_reference_source.py holds this
  oracle to those vectors: wrapped phase, table index and lifter index bit for bit; F0, WaveNet output, sub-bands, excitation,
  |VTF| and waveform to 1e-5 .. 1e-4 of peak.  What remains assumed is TensorFlow's arithmetic *inside* a primitive (summation
  order of a convolution, sequential float32 cumsum, SAME-padding split) -- SURVEY.md A.1.  ``make_reference_model_goldens.py``
  goes furthest: the reference's ``MBExWN`` object is created by its own constructors from this package's config.yaml and run
  through ``PaNWaveNet.infer``.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from mbexwn_vocoder_b200 import dsp_init

LOG_TO_DB = 20 * np.log10(np.exp(1))


# --------------------------------------------------------------------------------------------------
# layer restatements
# --------------------------------------------------------------------------------------------------

def weight_norm_kernel(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """conv_layers.py:149-153: kernel = g * l2_normalize(v, axis=[0, 1]) (tf epsilon 1e-12 on the squared norm)."""
    sq = torch.sum(v * v, dim=(0, 1), keepdim=True)
    return g * (v * torch.rsqrt(torch.clamp(sq, min=1e-12)))


def conv1d_keras(x: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor, padding: str, dilation: int = 1):
    """Keras Conv1D, channels-last, cross-correlation (conv_layers.py:154).  x (B, T, Cin), kernel (k, Cin, Cout)."""
    k = kernel.shape[0]
    xt = x.transpose(1, 2)
    if padding == "SAME":
        total = (k - 1) * dilation
        xt = F.pad(xt, (total // 2, total - total // 2))
    elif padding == "CAUSAL":                                 # Keras padding="causal": all (k - 1) d zeros on the left
        xt = F.pad(xt, ((k - 1) * dilation, 0))
    elif padding != "VALID":
        raise NotImplementedError(padding)
    y = F.conv1d(xt, kernel.permute(2, 1, 0).contiguous(), bias, dilation=dilation)
    return y.transpose(1, 2)


def pad1d(x: torch.Tensor, left: int, right: int, mode: str) -> torch.Tensor:
    """custom_layers.py:47-71.  SYMMETRIC mirrors including the edge sample, EDGE repeats it."""
    if left == 0 and right == 0:
        return x
    if mode == "EDGE":
        return torch.cat((x[:, :1].expand(-1, left, -1), x, x[:, -1:].expand(-1, right, -1)), dim=1)
    if mode == "SYMMETRIC":
        lhs = torch.flip(x[:, :left], dims=(1,))
        rhs = torch.flip(x[:, x.shape[1] - right:], dims=(1,)) if right else x[:, :0]
        return torch.cat((lhs, x, rhs), dim=1)
    if mode == "CONSTANT":
        return F.pad(x, (0, 0, left, right))
    raise NotImplementedError(mode)


def lin_interp(x: torch.Tensor, up: int, num_pad_end: int = 1, drop_last: bool = True) -> torch.Tensor:
    """support_layers.py:99-121 with the weights of :19-27 (computed in float64, cast to the layer dtype).

    depthwise_conv2d SAME with a width-2 kernel pads one zero on the right only.
    """
    if num_pad_end > 0:
        x = torch.cat((x, x[:, -1:].expand(-1, num_pad_end, -1)), dim=1)
    B, T, C = x.shape
    w0 = torch.tensor((up - np.arange(up)) / up, dtype=x.dtype)
    w1 = torch.tensor(np.arange(up) / up, dtype=x.dtype)
    nxt = torch.cat((x[:, 1:], torch.zeros_like(x[:, :1])), dim=1)
    res = x[:, :, None, :] * w0[None, None, :, None] + nxt[:, :, None, :] * w1[None, None, :, None]
    res = res.reshape(B, T * up, C)
    return res[:, :(T - 1) * up + (0 if drop_last else 1)]


def soft_sigmoid(x):
    """custom_AE_layers.py:91-99."""
    return 0.5 + 0.5 * x / (1 + torch.abs(x))


class SubNet:
    """Layer list of generate_subnet_from_specs (custom_pulsed_generator.py:38-148), inference only."""

    def __init__(self, specs, base_name, weights, dtype, final_n_channels, final_nks, final_activation,
                 target_ups=None, pad_to_valid=False, remove_inactive_pad_layers=False, use_prelu=True, alpha=0.2,
                 force_causal=False):
        self.layers: List[Tuple] = []
        default_padding = "CAUSAL" if force_causal else "SAME"            # custom_pulsed_generator.py:53

        def pads(ks):
            """TFPad1d sizes; force_causal moves all of them to the left (custom_pulsed_generator.py:76-81)."""
            pl, pr = (ks - 1) // 2 + ((ks - 1) % 2), (ks - 1) // 2
            return (pl + pr, 0, pl) if force_causal else (pl, pr, pl)
        self.w = weights
        self.dtype = dtype
        total_ups = 1
        if not specs:
            self.total_ups = total_ups
            return
        for ii, spec in enumerate(specs):
            if spec[0] == "L":
                self.layers.append(("lin", int(spec[1])))
                continue
            ks, nf = spec[0], spec[1]
            linear_up, up = False, 1
            if len(spec) > 2:
                if isinstance(spec[2], str):
                    if spec[2][0] == "L":
                        linear_up = True
                    up = int(spec[2][1:])
                else:
                    up = spec[2]
            pl, pr, active = pads(ks)
            name = f"{base_name}_Layer_{ii}"
            if linear_up:
                if (not remove_inactive_pad_layers) or active > 0:
                    self.layers.append(("pad", pl, pr, "EDGE" if pad_to_valid else "SYMMETRIC"))
                self.layers.append(("conv", name, "VALID", 1))
                self.layers.append(("lin", up))
            elif up > 1:
                if pad_to_valid and active > 0:
                    self.layers.append(("pad", pl, pr, "EDGE"))
                self.layers.append(("conv", name, "VALID" if pad_to_valid else default_padding, up))
            else:
                if (not remove_inactive_pad_layers) or active > 0:
                    self.layers.append(("pad", pl, pr, "EDGE" if pad_to_valid else "SYMMETRIC"))
                self.layers.append(("conv", name, "VALID", 1))
            self.layers.append(("prelu", f"{base_name}_ActLayer_{ii}") if use_prelu else ("leaky", alpha))
            total_ups *= up
        if final_nks is not None:
            pl, pr, active = pads(final_nks)
            if pad_to_valid and active > 0:
                self.layers.append(("pad", pl, pr, "EDGE"))
            self.layers.append(("conv", f"{base_name}_Layer_final", "VALID" if pad_to_valid else default_padding, 1))
            if target_ups is not None and total_ups != target_ups:
                up = target_ups // total_ups
                if total_ups * up != target_ups:
                    raise RuntimeError("get_missing_upsamling_factor::error")
                self.layers.append(("lin", up))
                total_ups *= up
            if final_activation is not None:
                self.layers.append(("final_act", final_activation))
        self.total_ups = total_ups

    def _t(self, name):
        return torch.as_tensor(self.w[name], dtype=self.dtype)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers:
            kind = layer[0]
            if kind == "pad":
                x = pad1d(x, layer[1], layer[2], layer[3])
            elif kind == "conv":
                _, name, padding, up = layer
                kern = weight_norm_kernel(self._t(f"{name}/v"), self._t(f"{name}/g"))
                x = conv1d_keras(x, kern, self._t(f"{name}/bias"), padding)
                if up > 1:                                    # conv_layers.py:250-255 depth-to-time
                    x = x.reshape(x.shape[0], x.shape[1] * up, -1)
            elif kind == "lin":
                x = lin_interp(x, layer[1])
            elif kind == "prelu":
                a = self._t(f"{layer[1]}/alpha")
                x = torch.relu(x) - a * torch.relu(-x)
            elif kind == "leaky":
                x = torch.where(x >= 0, x, layer[1] * x)
            elif kind == "final_act":
                if layer[1] == "soft_sigmoid":
                    x = soft_sigmoid(x)
                else:
                    raise NotImplementedError(layer[1])
        return x


# --------------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------------

class OracleMBExWN:
    """Inference branch of PaNWaveNet.infer + MBExWN.call (wavegen_1d.py:483-526, custom_pulsed_generator.py:556-771)."""

    def __init__(self, hparams: Dict, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.dtype = dtype
        self.np_dtype = np.float32 if dtype == torch.float32 else np.float64
        self.w = weights
        mc = copy.deepcopy(hparams["mbexwn_config"])
        pc = hparams["preprocess_config"]
        self.sample_rate = pc["sample_rate"]
        self.hop = pc["hop_size"]
        self.mel_channels = pc["mel_channels"]
        mb = mc["multi_band_config"]
        self.mb_factor = mb["subbands"]
        self.pulse_rate_factor = mc.get("pulse_rate_factor", 2)
        self.pulse_rate = self.sample_rate / self.pulse_rate_factor
        self.pulse_channels = mc.get("pulse_channels", 8)
        ups = list(mc["pp_mod_subnet_upsampling_factors"])
        self.sub_per_frame = self.hop // self.mb_factor
        self.pulse_per_frame = (self.sub_per_frame * self.pulse_channels) // int(np.prod(ups))
        self.f0_down = int(self.sample_rate // self.pulse_rate)
        self.fmin = mc.get("pp_min_frequency", 40.0)
        self.fmax = mc.get("pp_max_frequency", 600.0)
        self.sigma = mc.get("pp_mod_subnet_noise_channel_sigma", 0.5)
        use_prelu, alpha = mc.get("use_prelu", True), mc.get("alpha", 0.2)
        self.n_ceps = mc.get("ps_max_ceps_coefs", 120)
        self.env_order_scale = mc.get("ps_env_order_scale")
        fdb = mc.get("filter_max_db_range")
        self.filter_max_log_range = fdb / LOG_TO_DB if fdb is not None else None

        self.pp = SubNet(mc["pp_subnet"], "PulsPar", weights, dtype, 1, 1, mc.get("pp_activation", "soft_sigmoid"),
                         target_ups=self.pulse_per_frame, pad_to_valid=mc.get("pp_subnet_use_valid_padding", False),
                         remove_inactive_pad_layers=mc.get("remove_inactive_pad_layers", False),
                         use_prelu=use_prelu, alpha=alpha, force_causal=bool(mc.get("force_causal", False)))
        self.wn_padding = "CAUSAL" if mc.get("force_causal") else "SAME"       # custom_pulsed_generator.py:474-475
        self.ps_use_stft = bool(mc.get("ps_use_stft", True))
        self.ps_off = bool(mc.get("ps_off", False))
        self.preserve_energy = bool(mc.get("spect_filters_preserve_energy", False))
        # final width of the PS sub-net: cepstrum, or one log gain per sub-band (custom_pulsed_generator.py:422)
        ps_final = self.n_ceps if self.ps_use_stft else self.mb_factor
        self.ps = None if self.ps_off else SubNet(mc["ps_subnet"], "PS", weights, dtype, ps_final, 1, None,
                         pad_to_valid=mc.get("ps_subnet_use_valid_padding", False),
                         remove_inactive_pad_layers=mc.get("remove_inactive_pad_layers", False),
                         use_prelu=use_prelu, alpha=alpha, force_causal=bool(mc.get("force_causal", False)))

        wn = copy.deepcopy(mc["pp_mod_subnet"])
        n_channels = wn.pop("n_channels")
        self.cond_lin = wn.pop("cond_lin_upsampling", 16)
        self.cond_k = wn.pop("cond_kernel_size", 3)
        wn_rate = self.pulse_rate / self.pulse_channels
        spect_rate = self.sample_rate / self.hop
        # pp_waveNetBlocks (custom_pulsed_generator.py:459-488): one WaveNetAEBlock per (up-sampling factor, channel factor);
        # a block's conditioning conv up-samples to its own rate / cond_lin, its output rate is `up` times its input rate
        self.blocks = []
        rate = wn_rate
        for iwn, (u, cf) in enumerate(zip(ups, mc["pp_mod_subnet_channel_factors"])):
            self.blocks.append({"name": f"PP_waveNetBlock_ups{u}_{iwn}_WNBlock_WN", "C": int(n_channels * cf), "up": int(u),
                                "up_name": f"PP_waveNetBlock_ups{u}_{iwn}_WNBlock_UP_{u}",
                                "cond_conv_up": int(rate // (spect_rate * self.cond_lin))})
            rate *= u
        self.C = self.blocks[0]["C"]
        self.cond_conv_up = self.blocks[0]["cond_conv_up"]
        self.n_layers = wn.get("n_layers", 12)
        self.k = wn.get("kernel_size", 3)
        step, max_log2 = wn.get("dilation_rate_step", 1), wn.get("max_log2_dilation_rate", None)
        self.dilations = [2 ** (int(i // step) % max_log2) if max_log2 is not None else 2 ** int(i // step)
                          for i in range(self.n_layers)]            # custom_AE_layers.py:229-233
        self.gate = wn.get("activation", "gtu")
        self.wn_name = self.blocks[0]["name"]
        self.post_name = "MBExWNGen_PaNMPulseWaveNet_Post"

        self.win_size, self.fft_size = dsp_init.stft_sizes(self.sample_rate, self.hop, mc.get("internal_win_size_s"),
                                                           int(mc.get("internal_fft_over", 0)))
        self.wt = dsp_init.build_wavetables(sample_rate=self.pulse_rate, **mc["wavetable_config"])
        self.subharm = int(mc["wavetable_config"].get("add_subharm_chans", 0) or 0)
        self.pulse_pqmf = None
        if mc.get("pulse_channels_use_pqmf"):                                  # custom_pulsed_generator.py:499-501
            pq = mc["pulse_channels_multi_band_config"]
            self.pulse_pqmf = (dsp_init.pqmf_filters(pq["subbands"], pq["taps"], pq["cutoff_ratio"], pq["beta"])[0], int(pq["taps"]))
        self.taps = mb["taps"]
        _, self.h_syn = dsp_init.pqmf_filters(mb["subbands"], mb["taps"], mb["cutoff_ratio"], mb["beta"])
        self.window = dsp_init.hann_periodic(self.win_size)
        self.inv_window = dsp_init.inverse_stft_window(self.win_size, self.hop)
        self.smooth_kernel = dsp_init.f0_smoothing_kernel(self.hop)
        if self.env_order_scale:
            self.lifter_log10f0, self.lifters = dsp_init.cepstral_lifters(
                self.env_order_scale, self.sample_rate, self.fmin, self.fmax, self.n_ceps)

    # ---- helpers ---------------------------------------------------------------------------------
    def _t(self, name):
        return torch.as_tensor(self.w[name], dtype=self.dtype)

    def _conv(self, x, name, padding="SAME", dilation=1):
        kern = weight_norm_kernel(self._t(f"{name}/v"), self._t(f"{name}/g"))
        return conv1d_keras(x, kern, self._t(f"{name}/bias"), padding, dilation)

    # ---- stage 1: F0 -----------------------------------------------------------------------------
    def generate_f0(self, mel: torch.Tensor) -> torch.Tensor:
        """custom_pulsed_generator.py:773-791."""
        x = self.pp(mel)
        f0 = x[:, :, 0] * (self.fmax - self.fmin) + self.fmin
        return f0[:, :mel.shape[1] * self.pulse_per_frame]

    # ---- stage 2: pulse wavetable ------------------------------------------------------------------
    def stable_cumsum_and_wrap(self, v: np.ndarray, chunk_size: int = 1000) -> np.ndarray:
        """tf_wavetable.py:429-492; sequential accumulation in the working dtype, floor-mod 1."""
        dt = v.dtype
        B, n_time = v.shape
        rem = n_time % chunk_size
        if rem:
            v = np.concatenate((v, np.zeros((B, chunk_size - rem), dtype=dt)), axis=1)
        n_chunks = v.shape[1] // chunk_size
        phase = np.cumsum(v.reshape(B, n_chunks, chunk_size), axis=2, dtype=dt)
        offsets = np.mod(phase[:, :, -1:], dt.type(1))
        offsets = np.concatenate((np.zeros((B, 1, 1), dtype=dt), offsets), axis=1)[:, :-1]
        offsets = np.mod(np.cumsum(offsets, axis=1, dtype=dt), dt.type(1))
        phase = np.mod(phase + offsets, dt.type(1))
        return phase.reshape(B, -1)[:, :n_time]

    def pulse_generator(self, f0: np.ndarray) -> Dict[str, np.ndarray]:
        """PulseWaveTable.call + _linear_lookup (tf_wavetable.py:495-560, :605-638)."""
        dt = self.np_dtype
        f0 = f0.astype(dt)
        tab = self.wt.tables.astype(dt)
        v = f0 / dt(self.pulse_rate)
        phase = self.stable_cumsum_and_wrap(v)
        p = phase * dt(self.wt.n_period)
        pq = np.floor(p)
        frac = p - pq
        i0 = pq.astype(np.int32)
        one_m = dt(1.0) - frac
        samples = tab[i0] * one_m[:, :, None] + tab[i0 + 1] * frac[:, :, None]          # (B, N, K)
        ratio = np.maximum(dt(self.wt.min_transposition),
                           np.minimum(dt(self.wt.max_transposition), f0 / dt(np.float32(self.wt.nominal_f0))))
        diff = np.log(ratio)[:, :, None] * dt(self.wt.grid_norm) - np.arange(tab.shape[1], dtype=dt)
        mix = np.maximum(dt(1) - np.abs(diff), dt(0))
        pulse = np.sum(samples * mix, axis=2, dtype=dt)
        return {"phase": phase, "index": i0, "frac": frac, "pulse": pulse}

    # ---- stage 3/4: conditioning + WaveNet ---------------------------------------------------------
    def conditioning(self, mel: torch.Tensor, block: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """cond_ sub-pixel conv + LinInterp (custom_AE_layers.py:215-227, :287-289)."""
        blk = self.blocks[block]
        c = self._conv(mel, f"{blk['name']}/cond_", self.wn_padding)
        c = c.reshape(c.shape[0], c.shape[1] * blk["cond_conv_up"], -1)
        return c, lin_interp(c, self.cond_lin)

    def wavenet(self, x: torch.Tensor, mel: torch.Tensor, taps: Optional[Dict] = None) -> torch.Tensor:
        """The pp_waveNetBlocks loop (custom_pulsed_generator.py:908-910): WaveNetAEBlock.call (custom_AE_layers.py:564-575) =
        WaveNetAE, then the sub-pixel up-sampling conv (k = 3, conv_layers.py:250-255) when the block's factor is > 1."""
        for ib, blk in enumerate(self.blocks):
            x = self.wavenet_block(x, mel, ib, taps if ib == len(self.blocks) - 1 else None)
            if blk["up"] > 1:
                y = self._conv(x, blk["up_name"], self.wn_padding)
                x = y.reshape(y.shape[0], y.shape[1] * blk["up"], -1)
            if taps is not None:
                taps[f"block_out_{ib}"] = x
        return x

    def wavenet_block(self, x: torch.Tensor, mel: torch.Tensor, block: int = 0, taps: Optional[Dict] = None) -> torch.Tensor:
        """WaveNetAE.call (custom_AE_layers.py:273-346), n_ch_groups = 1, shared up-sampled conditioning."""
        n = self.blocks[block]["name"]
        h = self._conv(x, f"{n}/start")
        cond_lo, cond = self.conditioning(mel, block)
        if taps is not None:
            taps["cond_lo"], taps["h0"] = cond_lo, h
        out = None
        for i, d in enumerate(self.dilations):
            z = self._conv(h, f"{n}/conv1D_{i}", self.wn_padding, d) + cond
            a, b = torch.split(z, z.shape[-1] // 2, dim=-1)
            if self.gate == "gtu":
                a = torch.tanh(a)
            elif self.gate == "gfu":
                a = a / (1 + torch.abs(a))
            elif self.gate == "gsu":
                a = a / (1 + torch.sqrt(torch.abs(a)))
            act = a * torch.sigmoid(b)
            rs = self._conv(act, f"{n}/res_skip_{i}")
            if i < self.n_layers - 1:
                res, skip = torch.split(rs, rs.shape[-1] // 2, dim=-1)
                h = h + res
            else:
                skip = rs
            out = skip if out is None else out + skip
            if taps is not None:
                taps[f"act_{i}"], taps[f"h_{i + 1}"] = act, h
        wn_out = self._conv(out, f"{n}/end")
        if taps is not None:
            taps["skip"] = out
            taps["wn_out"] = wn_out
        return wn_out

    def pqmf_synthesis(self, x: torch.Tensor) -> torch.Tensor:
        """TFPQMF.synthesis (tf_preprocess.py:208-226): zero-stuff x S with gain S, pad taps/2, correlate, sum bands."""
        B, T, S = x.shape
        up = torch.zeros(B, T * S, S, dtype=x.dtype)
        up[:, ::S] = x * S
        up = F.pad(up.transpose(1, 2), (self.taps // 2, self.taps // 2))
        kern = torch.as_tensor(self.h_syn, dtype=x.dtype)[None]                 # (1, S, taps+1)
        return F.conv1d(up, kern)[:, 0]

    def generate_multiband_gain(self, mel: torch.Tensor) -> torch.Tensor:
        """custom_pulsed_generator.py:857-884 (ps_use_stft = False): exp of the PS sub-net's per-band log gain."""
        x = self.ps(mel)
        if self.preserve_energy:
            x = x - torch.mean(x, dim=-1, keepdim=True)
        return torch.exp(x)

    def generate_excitation(self, mel: torch.Tensor, f0: torch.Tensor, noise: torch.Tensor,
                            taps: Optional[Dict] = None, mb_gain: Optional[torch.Tensor] = None) -> torch.Tensor:
        """custom_pulsed_generator.py:886-925.  `noise` is the N(0,1) draw of :906, shape (B, 20T, 1)."""
        pg = self.pulse_generator(f0.detach().cpu().numpy())
        pulse = torch.as_tensor(pg["pulse"], dtype=self.dtype)
        if self.pulse_pqmf is not None:
            # TFPQMF.analysis (tf_preprocess.py:192-202): zero pad taps/2, cross-correlate with the analysis bank, keep every
            # S-th sample; sub-harmonic channels are appended folded (custom_pulsed_generator.py:895-900)
            ana, taps_p = self.pulse_pqmf
            S = ana.shape[0]
            xp = F.pad(pulse[:, None, :], (taps_p // 2, taps_p // 2))
            y = F.conv1d(xp, torch.as_tensor(ana, dtype=self.dtype)[:, None, :], stride=1)      # (B, S, N)
            x = y[:, :, ::S].transpose(1, 2)                                                   # (B, N / S, S)
            if self.subharm:
                dt = self.np_dtype
                w2pi = pg["phase"].astype(dt) * dt(2) * dt(np.float32(np.pi))
                sub = np.stack([np.sin(w2pi / dt(ii)) for ii in range(2, self.subharm + 2)], axis=-1)
                x = torch.cat((x, torch.as_tensor(sub, dtype=self.dtype).reshape(pulse.shape[0], -1,
                                                                                self.pulse_channels * self.subharm)), dim=-1)
        elif self.subharm:                                                     # tf_wavetable.py:520-521, :554-559
            dt = self.np_dtype
            w2pi = pg["phase"].astype(dt) * dt(2) * dt(np.float32(np.pi))
            chans = [pg["pulse"].astype(dt)] + [np.sin(w2pi / dt(ii)) for ii in range(2, self.subharm + 2)]
            pulse_all = torch.as_tensor(np.stack(chans, axis=-1), dtype=self.dtype)
            x = pulse_all.reshape(pulse.shape[0], -1, self.pulse_channels * (1 + self.subharm))   # custom_pulsed_generator.py:893
        else:
            x = pulse.reshape(pulse.shape[0], -1, self.pulse_channels)
        if self.sigma:
            x = torch.cat((x, self.sigma * noise.to(self.dtype)), dim=-1)
        if taps is not None:
            taps.update({"phase": pg["phase"], "index": pg["index"], "frac": pg["frac"], "pulse": pulse, "wn_in": x})
        y = self.wavenet(x, mel, taps)
        sub = self._conv(y, self.post_name)
        if mb_gain is not None:                                               # :916-917
            sub = sub * mb_gain[:, :sub.shape[1]]
        exc = self.pqmf_synthesis(sub)
        if taps is not None:
            taps.update({"subbands": sub, "excitation": exc})      # "wn_out" = the last block's WaveNetAE output
        return exc

    # ---- stage 5: vocal-tract filter ---------------------------------------------------------------
    def cepstral_windows(self, f0: torch.Tensor) -> torch.Tensor:
        """_get_cepstral_windows (custom_pulsed_generator.py:507-525)."""
        kern = torch.as_tensor(self.smooth_kernel, dtype=self.dtype)
        half = kern.shape[0] // 2
        padded = torch.cat((f0[:, :1].expand(-1, half), f0, f0[:, -1:].expand(-1, half)), dim=1)
        sm = F.conv1d(padded[:, None], kern[None, None], stride=self.pulse_per_frame)[:, 0]
        grid = torch.as_tensor(self.lifter_log10f0, dtype=self.dtype)
        log10 = torch.clamp((1 / np.log(10)) * torch.log(sm), min=grid[0], max=grid[-1])
        ratio = (log10 - grid[0]) / (grid[-1] - grid[0])
        idx = torch.round(ratio * (grid.shape[0] - 1)).to(torch.int64)         # round-half-even like tf.round
        return torch.as_tensor(self.lifters, dtype=self.dtype)[idx], idx

    def generate_specenv(self, mel: torch.Tensor, f0: torch.Tensor, taps: Optional[Dict] = None) -> torch.Tensor:
        """custom_pulsed_generator.py:793-855 with spect_filters_preserve_energy = False."""
        ceps = self.ps(mel)
        if taps is not None:
            taps["ceps"] = ceps
        if self.env_order_scale:
            win, idx = self.cepstral_windows(f0)
            assert bool(torch.all(win[:, :, 0] == 1.0)), "problems with generated cepstral windows"   # :807
            ceps = ceps * win
            if taps is not None:
                taps["lifter_index"] = idx
        full = F.pad(ceps[:, :, 1:], (1, self.fft_size - ceps.shape[2]))
        log_spec = torch.fft.rfft(full)
        if self.filter_max_log_range:
            vtf = torch.exp(torch.complex(self.filter_max_log_range * torch.tanh(log_spec.real), log_spec.imag))
        else:
            vtf = torch.exp(log_spec)
        if taps is not None:
            taps["vtf"] = vtf
        return vtf

    def stft_filter(self, exc: torch.Tensor, vtf: torch.Tensor, n_frames: int, n_pulse: int) -> torch.Tensor:
        """custom_pulsed_generator.py:681-724: pad, STFT (periodic Hann, rfft), multiply, inverse STFT, crop."""
        half = self.win_size // 2
        padded = F.pad(exc, (half, half + self.hop + 1))
        frames = padded.unfold(-1, self.win_size, self.hop)[:, :n_frames]
        spec = torch.fft.rfft(frames * torch.as_tensor(self.window, dtype=self.dtype), n=self.fft_size)
        sig = torch.fft.irfft(spec * vtf, n=self.fft_size)[..., :self.win_size]
        sig = sig * torch.as_tensor(self.inv_window, dtype=self.dtype)
        B = sig.shape[0]
        out = torch.zeros(B, (n_frames - 1) * self.hop + self.win_size, dtype=self.dtype)
        for f in range(n_frames):                                            # tf.signal.overlap_and_add
            out[:, f * self.hop:f * self.hop + self.win_size] += sig[:, f]
        return out[:, half:half + n_pulse * self.f0_down]

    # ---- whole forward -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, mel: np.ndarray, noise: np.ndarray, f0_override: Optional[np.ndarray] = None,
                return_taps: bool = True) -> Dict[str, np.ndarray]:
        """mel (B, T, n_mel) scaled log-mel; noise (B, T*steps, 1) standard normal.  Returns taps incl. 'waveform'."""
        mel_t = torch.as_tensor(mel, dtype=self.dtype)
        taps: Dict = {}
        synth_length = mel_t.shape[1] * self.hop                               # mel_inverter.py:152
        f0 = self.generate_f0(mel_t) if f0_override is None else torch.as_tensor(f0_override, dtype=self.dtype)
        taps["F0"] = f0
        if (not self.ps_use_stft) or self.ps_off:                              # custom_pulsed_generator.py:666-674
            gain = None
            if not self.ps_off:
                # ps_gain_interpolator: LinInterp x hop, num_pad_end=1, drop_last=False (:453); the gain is produced at
                # the sample rate but applied to the sub-band rows, which read its first T * steps values (:917)
                gain = lin_interp(self.generate_multiband_gain(mel_t), self.hop, num_pad_end=1, drop_last=False)
                taps["mb_gain"] = gain[:, :mel_t.shape[1] * self.sub_per_frame]
            sig = self.generate_excitation(mel_t, f0, torch.as_tensor(noise), taps, mb_gain=gain)
            taps["waveform"] = sig[:, :synth_length]
            if not return_taps:
                return {"waveform": taps["waveform"].numpy()}
            return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in taps.items()}
        exc = self.generate_excitation(mel_t, f0, torch.as_tensor(noise), taps)
        vtf = self.generate_specenv(mel_t, f0, taps)
        sig = self.stft_filter(exc, vtf, mel_t.shape[1], f0.shape[1])
        taps["waveform"] = sig[:, :synth_length]                               # wavegen_1d.py:504-512
        if not return_taps:
            return {"waveform": taps["waveform"].numpy()}
        return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in taps.items()}


def synthetic_mel(n_frames: int, utt_id: int = 0, n_mel: int = 80) -> np.ndarray:
    """Synthetic scaled log-mel of SURVEY.md 8d: clip(N(-4, 2), log 1e-5, 2), 5-frame box smoothing in time."""
    g = torch.Generator().manual_seed(1234 + utt_id)
    x = torch.randn(n_frames + 4, n_mel, generator=g) * 2.0 - 4.0
    x = torch.clamp(x, min=float(np.log(1e-5)), max=2.0)
    x = x.unfold(0, 5, 1).mean(dim=-1)
    return x.numpy().astype(np.float32)


def synthetic_noise(n_steps: int, utt_id: int = 0) -> np.ndarray:
    g = torch.Generator().manual_seed(4321 + utt_id)
    return torch.randn(n_steps, 1, generator=g).numpy().astype(np.float32)
