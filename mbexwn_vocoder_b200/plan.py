"""Config -> execution plan for the MBExWN forward path.

Turns the ``mbexwn_config`` / ``preprocess_config`` dictionaries (the keyword arguments of the reference's
``MBExWN.__init__``, custom_pulsed_generator.py:155-504) into the flat description the CUDA engine
consumes: internal rates, the two mel-rate conv sub-nets as op lists, the WaveNet geometry, and the
init-time DSP constants.  This is the host-side counterpart of the reference's layer construction
(``generate_subnet_from_specs`` custom_pulsed_generator.py:38-148, ``WaveNetAE.__init__``
custom_AE_layers.py:120-265); nothing here runs per utterance.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import dsp_init

PAD_ZERO, PAD_SYMMETRIC, PAD_EDGE = 0, 1, 2
PS_STFT, PS_BAND_GAIN, PS_OFF = 0, 1, 2
ACT_NONE, ACT_PRELU, ACT_LEAKY, ACT_SOFT_SIGMOID_AFFINE, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4, 5
GATE_GTU, GATE_GLU, GATE_GFU, GATE_GSU = 0, 1, 2, 3
_GATES = {"gtu": GATE_GTU, "glu": GATE_GLU, "gfu": GATE_GFU, "gsu": GATE_GSU}


@dataclass
class ConvLayer:
    """One weight-normalised Conv1D (conv_layers.py:21-174) as the engine sees it."""
    name: str                   # reference layer name; weights are <name>/v, <name>/g, <name>/bias
    k: int
    cin: int
    cout: int                   # conv output channels *before* the sub-pixel unfold
    dilation: int = 1
    pad_l: int = 0
    pad_r: int = 0
    pad_mode: int = PAD_ZERO
    subpixel: int = 1           # time unfold factor: (T, cout) is re-read as (T*f, cout/f)
    init_std: Optional[float] = None     # RandomNormal stddev, None => Glorot uniform (Keras default)
    cb_free: int = 0            # checkerboard-free init factor (conv_layers.py:73-77)


@dataclass
class Op:
    """One step of a mel-rate sub-net program."""
    kind: str                   # "conv" | "lininterp"
    conv: Optional[ConvLayer] = None
    up: int = 1                 # lininterp factor
    act: int = ACT_NONE         # activation fused after this op
    act_name: Optional[str] = None       # reference name of the activation layer (PReLU alpha lives there)
    act_channels: int = 0
    rate_in: int = 1            # rows per mel frame at the op input
    rate_out: int = 1           # rows per mel frame at the op output
    ch_out: int = 0             # channels per row at the op output


@dataclass
class WaveNetSpec:
    name: str
    c: int                      # residual channels
    c_in: int                   # input channels (pulse_channels [+1 noise])
    c_out: int                  # n_out_channels of `end`
    n_layers: int
    k: int
    dilations: List[int]
    gate: int
    cond_k: int
    cond_conv_up: int           # sub-pixel factor of the conditioning conv
    cond_lin_up: int            # linear interpolation factor after it
    steps_per_frame: int        # WaveNet rows per mel frame
    cond_cin: int = 80          # mel channels feeding the conditioning conv
    causal: bool = False        # force_causal: padding "CAUSAL" for the dilated and the conditioning convs
    up: int = 1                 # WaveNetAEBlock.up_down_factor: sub-pixel up-sampling conv after the block (1 = none)

    @property
    def up_name(self) -> str:
        """Layer name of the block's TF2C_Conv1DUpDownSample (custom_AE_layers.py:523-526)."""
        return f"{self.name}_WNBlock_UP_{self.up}"


@dataclass
class NormMelSpec:
    """NormMelComponents constructor arguments that matter at inference (wavegen_1d.py:580-632)."""
    iters: int
    win: int
    smooth_win: int
    squared_win: bool
    use_pinv: bool
    norm_fact: float
    floor: float                # 1 / max_norm_fact, 0 = none
    compress_exp: float         # 0 = none
    lin_scale: float
    lin_off: float
    mel_scale: float
    use_max_limit: bool
    sample_rate: int
    fft_size: int
    n_mel: int
    fmin: float
    fmax: float
    # constants (finalize_plan)
    proj: Optional[np.ndarray] = None
    proj_cols: int = 0
    proj_scale: float = 1.0
    smooth_window: Optional[np.ndarray] = None
    gwin: Optional[np.ndarray] = None


@dataclass
class ModelPlan:
    sample_rate: int
    hop: int
    mel_channels: int
    pulse_rate_factor: int
    pulse_rate: float
    pulse_channels: int
    subbands: int
    pulse_per_frame: int        # spect_to_pulse_upsampling_factor
    steps_per_frame: int
    f0_min: float
    f0_max: float
    noise_sigma: float
    alpha: float
    use_prelu: bool
    pp_ops: List[Op]
    ps_ops: List[Op]
    n_ceps: int
    wavenet: WaveNetSpec
    post_name: str
    pqmf_cfg: Dict
    stft_win: int
    fft_size: int
    filter_max_log_range: Optional[float]
    env_order_scale: Optional[float]
    wavetable_cfg: Dict
    max_halo_frames: int = 1
    subharm: int = 0            # wavetable_config.add_subharm_chans
    pulse_pqmf_cfg: Optional[Dict] = None    # pulse_channels_multi_band_config when pulse_channels_use_pqmf
    pulse_pqmf_ana: Optional[np.ndarray] = None   # (pulse_channels, taps + 1) analysis bank (finalize)
    ps_mode: int = PS_STFT      # PS_STFT | PS_BAND_GAIN | PS_OFF
    ps_preserve_energy: bool = False
    norm: Optional["NormMelSpec"] = None     # NormMelComponents (normalize_rms_from_mell), None = off
    # pp_waveNetBlocks (custom_pulsed_generator.py:459-488): `wavenet` is block 0 (its rate = steps_per_frame, the rate of the
    # pulse / noise input); later blocks run `up` times faster each; the last one ends at sub_per_frame = hop / subbands
    wavenet_blocks: List["WaveNetSpec"] = field(default_factory=list)
    sub_per_frame: int = 0
    # init-time constants (filled by finalize())
    wavetables: Optional[dsp_init.WaveTables] = None
    pqmf_syn: Optional[np.ndarray] = None
    pqmf_poly: Optional[np.ndarray] = None
    pqmf_q: int = 0
    pqmf_back: int = 0
    window: Optional[np.ndarray] = None
    inv_window: Optional[np.ndarray] = None
    lifter_log10f0: Optional[np.ndarray] = None
    lifters: Optional[np.ndarray] = None
    f0_smooth: Optional[np.ndarray] = None

    def conv_layers(self) -> List[ConvLayer]:
        """Every weight-carrying conv in reference construction order."""
        out = [op.conv for op in self.pp_ops if op.kind == "conv"]
        out += [op.conv for op in self.ps_ops if op.kind == "conv"]
        for wn in self.blocks:
            out += wavenet_layers(wn)
        out.append(ConvLayer(self.post_name, 1, self.wavenet.c_out, self.subbands))
        return out

    @property
    def blocks(self) -> List["WaveNetSpec"]:
        return self.wavenet_blocks or [self.wavenet]


def _pad_sizes(ks: int) -> Tuple[int, int]:
    return (ks - 1) // 2 + ((ks - 1) % 2), (ks - 1) // 2


def _same_pad(ks: int, dilation: int = 1) -> Tuple[int, int]:
    total = (ks - 1) * dilation
    return total // 2, total - total // 2


def subnet_program(specs, base_name: str, cin: int, final_n_channels: int, final_nks: Optional[int],
                   final_act: int, init_std: float, target_ups: Optional[int], use_prelu: bool,
                   pad_to_valid: bool = False, remove_inactive_pad_layers: bool = False,
                   force_causal: bool = False) -> Tuple[List[Op], int]:
    """Op list of one conv sub-net; restates generate_subnet_from_specs (custom_pulsed_generator.py:38-148).

    Padding layers are folded into the conv that follows them; sub-pixel unfolds are free re-views.
    """
    act = ACT_PRELU if use_prelu else ACT_LEAKY
    if force_causal:                                    # every pad on the left (custom_pulsed_generator.py:53, :76-81)
        _pads = lambda ks: (ks - 1, 0)
        _same = lambda ks: (ks - 1, 0)
    else:
        _pads, _same = _pad_sizes, _same_pad
    pad_kind = PAD_EDGE if pad_to_valid else PAD_SYMMETRIC
    ops: List[Op] = []
    rate, ch, total_ups = 1, cin, 1
    if not specs:
        return ops, total_ups
    for ii, spec in enumerate(specs):
        if spec[0] == "L":
            up = int(spec[1])
            ops.append(Op("lininterp", up=up, rate_in=rate, rate_out=rate * up, ch_out=ch))
            rate *= up                        # quirk Q2: no activation, total_ups not updated
            continue
        ks, nf = int(spec[0]), int(spec[1])
        linear_up, up = False, 1
        if len(spec) > 2:
            if isinstance(spec[2], str):
                linear_up = spec[2][0] == "L"
                up = int(spec[2][1:])
            else:
                up = int(spec[2])
        pl, pr = _pads(ks)
        name, act_name = f"{base_name}_Layer_{ii}", f"{base_name}_ActLayer_{ii}"
        if linear_up:
            conv = ConvLayer(name, ks, ch, nf, pad_l=pl, pad_r=pr, pad_mode=pad_kind, init_std=init_std)
            ops.append(Op("conv", conv=conv, rate_in=rate, rate_out=rate, ch_out=nf))
            ops.append(Op("lininterp", up=up, act=act, act_name=act_name, act_channels=nf,
                          rate_in=rate, rate_out=rate * up, ch_out=nf))
        elif up > 1:
            if pad_to_valid:
                conv = ConvLayer(name, ks, ch, nf * up, pad_l=pl, pad_r=pr, pad_mode=PAD_EDGE, subpixel=up,
                                 init_std=init_std, cb_free=up)
            else:
                sl, sr = _same(ks)
                conv = ConvLayer(name, ks, ch, nf * up, pad_l=sl, pad_r=sr, pad_mode=PAD_ZERO, subpixel=up,
                                 init_std=init_std, cb_free=up)
            ops.append(Op("conv", conv=conv, act=act, act_name=act_name, act_channels=nf,
                          rate_in=rate, rate_out=rate * up, ch_out=nf))
        else:
            conv = ConvLayer(name, ks, ch, nf, pad_l=pl, pad_r=pr, pad_mode=pad_kind, init_std=init_std)
            ops.append(Op("conv", conv=conv, act=act, act_name=act_name, act_channels=nf,
                          rate_in=rate, rate_out=rate, ch_out=nf))
        rate *= up
        total_ups *= up
        ch = nf
    if final_nks is not None:
        if pad_to_valid:
            pl, pr = _pads(final_nks)
            conv = ConvLayer(f"{base_name}_Layer_final", final_nks, ch, final_n_channels, pad_l=pl, pad_r=pr,
                             pad_mode=PAD_EDGE, init_std=init_std)
        else:
            sl, sr = _same(final_nks)
            conv = ConvLayer(f"{base_name}_Layer_final", final_nks, ch, final_n_channels, pad_l=sl, pad_r=sr,
                             pad_mode=PAD_ZERO, init_std=init_std)
        ch = final_n_channels
        needs_up = target_ups is not None and total_ups != target_ups
        ops.append(Op("conv", conv=conv, act=ACT_NONE if needs_up else final_act, act_channels=ch,
                      rate_in=rate, rate_out=rate, ch_out=ch))
        if needs_up:
            up = target_ups // total_ups
            if total_ups * up != target_ups:
                raise RuntimeError(f"get_missing_upsamling_factor::error:: Upsampling to target upsampling factor "
                                   f"{target_ups} from {total_ups} is not possible for subnet {base_name}")
            ops.append(Op("lininterp", up=up, act=final_act, act_channels=ch,
                          rate_in=rate, rate_out=rate * up, ch_out=ch))
            rate *= up
            total_ups *= up
    return ops, total_ups


def wavenet_layers(wn: WaveNetSpec) -> List[ConvLayer]:
    """Weight-carrying layers of one WaveNetAE in construction order (custom_AE_layers.py:177-259)."""
    n = wn.name + "_WNBlock_WN"
    cl, cr = ((wn.cond_k - 1), 0) if wn.causal else _same_pad(wn.cond_k)
    layers = [ConvLayer(f"{n}/start", 1, wn.c_in, wn.c),
              ConvLayer(f"{n}/end", 1, wn.c, wn.c_out),
              ConvLayer(f"{n}/cond_", wn.cond_k, wn.cond_cin, 2 * wn.c * wn.cond_conv_up, pad_l=cl, pad_r=cr,
                        subpixel=wn.cond_conv_up, cb_free=wn.cond_conv_up)]
    for i, d in enumerate(wn.dilations):
        pl, pr = ((wn.k - 1) * d, 0) if wn.causal else _same_pad(wn.k, d)
        layers.append(ConvLayer(f"{n}/conv1D_{i}", wn.k, wn.c, 2 * wn.c, dilation=d, pad_l=pl, pad_r=pr))
        layers.append(ConvLayer(f"{n}/res_skip_{i}", 1, wn.c, 2 * wn.c if i < wn.n_layers - 1 else wn.c))
    if wn.up > 1:                                       # WaveNetAEBlock.up_down_sample (custom_AE_layers.py:519-526): k = 3
        ul, ur = (2, 0) if wn.causal else _same_pad(3)
        layers.append(ConvLayer(wn.up_name, 3, wn.c_out, wn.c_out * wn.up, pad_l=ul, pad_r=ur, subpixel=wn.up))
    return layers


def build_plan(hparams: Dict, finalize: bool = True) -> ModelPlan:
    """Validate a full config dict and derive the plan (MBExWN.__init__, custom_pulsed_generator.py:232-504)."""
    if "mbexwn_config" not in hparams:
        raise NotImplementedError(f"create_model::error::unkown config requested {list(hparams.keys())}. "
                                  f"Only mbexwn_config is currently supported.")        # models.py:31
    if not hparams.get("use_tf25_compatible_implementation", None):
        raise NotImplementedError("MBExWN::error::implmentations not selecting use_tf25_compatible_implementation "
                                  "are not supported")                                   # custom_pulsed_generator.py:272
    mc = copy.deepcopy(hparams["mbexwn_config"])
    pc = hparams["preprocess_config"]
    norm = None
    if mc.get("normalize_rms_from_mell", False):                                        # wavegen_1d.py:333-335
        win = int(pc.get("win_size", pc["fft_size"]))
        if 4 * int(pc["hop_size"]) != win:                                                # wavegen_1d.py:592-594
            raise RuntimeError(f"NormMelComponents:error: this module currently supports only the case where win_size "
                               f"{win} = 4 * hop_size {pc['hop_size']}")
        iters = int(mc.get("normalize_rms_num_smooth_iters", 0))
        if iters <= 0:
            raise NotImplementedError("normalize_rms_num_smooth_iters = 0 (per-channel time average, wavegen_1d.py:722) "
                                      "is not built")
        if int(mc.get("n_group", 1)) != 1:
            raise NotImplementedError("n_group != 1 is not built")
        cexp = mc.get("normalize_compressor_exp")
        norm = NormMelSpec(
            iters=iters, win=win, smooth_win=int(win * mc.get("normalize_smooth_win_scale", 1)),
            squared_win=bool(mc.get("normalize_smooth_with_squared_win", True)),
            use_pinv=bool(mc.get("normalize_use_pinv", False)), norm_fact=float(pc["fft_size"] * win * 0.5),
            floor=float(1. / mc["max_norm_fact"]) if mc.get("max_norm_fact") else 0.0,
            compress_exp=float(cexp) if cexp is not None else 0.0,
            lin_scale=float(mc.get("lin_amp_scale", 1.)), lin_off=float(mc.get("lin_amp_off", 1.e-5)),
            mel_scale=float(mc.get("mel_amp_scale", 1.)), use_max_limit=bool(mc.get("use_max_limit", False)),
            sample_rate=int(pc["sample_rate"]), fft_size=int(pc["fft_size"]), n_mel=int(pc["mel_channels"]),
            fmin=pc["fmin"], fmax=pc["fmax"])
        if cexp is not None and float(cexp) == 0.0:
            raise NotImplementedError("normalize_compressor_exp = 0 is not built")
    for key in ("normalize_rms_from_mell", "normalize_rms_num_smooth_iters", "normalize_compressor_exp",
                "normalize_smooth_win_scale", "normalize_smooth_with_squared_win", "normalize_use_pinv"):
        mc.pop(key, None)                                                               # wavegen_1d.py:416-422
    if "ps_max_db_range" in mc:                                                         # wavegen_1d.py:424-430
        mc["filter_max_db_range"] = mc.pop("ps_max_db_range")
        mc.pop("ns_max_db_range", None)
    if "pulse_rate_factor" not in mc:
        raise NotImplementedError("PaNWaveNet::error:: required parameter pulse_rate_factor is missing in your "
                                  "model config.")                                       # wavegen_1d.py:436

    sr, hop, n_mel = int(pc["sample_rate"]), int(pc["hop_size"]), int(pc["mel_channels"])
    mb = copy.deepcopy(mc["multi_band_config"])
    S = int(mb["subbands"])
    prf = int(mc.get("pulse_rate_factor", 2))
    pch = int(mc.get("pulse_channels", 8))
    ups_factors = list(mc["pp_mod_subnet_upsampling_factors"])
    chan_factors = list(mc["pp_mod_subnet_channel_factors"])
    pulse_rate = sr / prf
    if pulse_rate / pch * np.prod(ups_factors) * S != sr:                               # :344
        raise RuntimeError(f"MBExWN::config_error::the generated sample rate "
                           f"{pulse_rate / pch * np.prod(ups_factors) * S} != {sr}")
    # blocks = zip(upsampling factors, channel factors) (:465): the shorter list decides
    n_blocks = min(len(ups_factors), len(chan_factors))
    if n_blocks < 1 or any(int(u) != u or u < 1 for u in ups_factors):
        raise RuntimeError(f"MBExWN::config_error::pp_mod_subnet_upsampling_factors {ups_factors} / "
                           f"pp_mod_subnet_channel_factors {chan_factors}")
    if int(np.prod(ups_factors[:n_blocks])) != int(np.prod(ups_factors)):
        raise RuntimeError("MBExWN::config_error::up-sampling factors without a channel factor")
    ups_factors, chan_factors = [int(u) for u in ups_factors[:n_blocks]], list(chan_factors[:n_blocks])
    for key, why in (("pp_subnet_training_only", "pp_subnet_training_only"),):
        if mc.get(key):
            raise NotImplementedError(f"{why} is not built yet (SURVEY 8f-4)")
    # spectral shaping of the excitation (custom_pulsed_generator.py:666-724): STFT-domain vocal-tract filter (default),
    # per-band gain from the PS sub-net (ps_use_stft = False, :857-884), or none (ps_off)
    ps_mode = PS_OFF if mc.get("ps_off") else (PS_STFT if mc.get("ps_use_stft", True) else PS_BAND_GAIN)
    preserve_energy = bool(mc.get("spect_filters_preserve_energy", False))
    if preserve_energy and ps_mode == PS_STFT:
        raise NotImplementedError("spect_filters_preserve_energy with the STFT filter is not built yet (SURVEY 8f-4)")
    if not mc.get("pp_mod_subnet_use_pqmf", True):
        raise NotImplementedError("pp_mod_subnet_use_pqmf=False is not built yet")

    sub_per_frame = hop // S
    pulse_per_frame = (sub_per_frame * pch) // int(np.prod(ups_factors))               # :266
    use_prelu = bool(mc.get("use_prelu", True))
    pp_act = mc.get("pp_activation", "soft_sigmoid")
    if pp_act != "soft_sigmoid":
        raise NotImplementedError(f"pp_activation {pp_act} is not built yet")
    n_ceps = int(mc.get("ps_max_ceps_coefs", 120))

    pp_ops, _ = subnet_program(mc["pp_subnet"], "PulsPar", n_mel, 1, 1, ACT_SOFT_SIGMOID_AFFINE, 0.02,
                               pulse_per_frame, use_prelu, bool(mc.get("pp_subnet_use_valid_padding", False)),
                               bool(mc.get("remove_inactive_pad_layers", False)), bool(mc.get("force_causal", False)))
    if not pp_ops:
        raise NotImplementedError("models without pp_subnet (constant F0) are not built yet")
    ps_ops: List[Op] = []
    if ps_mode != PS_OFF:
        ps_ops, _ = subnet_program(mc["ps_subnet"], "PS", n_mel, n_ceps if ps_mode == PS_STFT else S, 1, ACT_NONE, 0.01,
                                   None, use_prelu, bool(mc.get("ps_subnet_use_valid_padding", False)),
                                   bool(mc.get("remove_inactive_pad_layers", False)),
                                   bool(mc.get("force_causal", False)))                # final width: :422
        if not ps_ops or ps_ops[-1].rate_out != 1:
            raise NotImplementedError("VTF sub-net must stay at mel-frame rate")

    wn_cfg = copy.deepcopy(mc["pp_mod_subnet"])
    n_channels = wn_cfg.pop("n_channels")
    cond_lin = int(wn_cfg.pop("cond_lin_upsampling", 16))
    cond_k = int(wn_cfg.pop("cond_kernel_size", 3))
    wn_rate = pulse_rate / pch
    spect_rate = sr / hop
    block_rates = []
    for u in ups_factors:                                                               # :465-488
        if wn_rate != (wn_rate // (spect_rate * cond_lin)) * spect_rate * cond_lin:     # :469
            raise RuntimeError(f"MBExWN::config_error:: cannot achieve conditioning rate {wn_rate} by means of integer "
                               f"usampling of spectrum rate {spect_rate} with linear up {cond_lin}")
        block_rates.append(wn_rate)
        wn_rate *= u
    wn_rate = block_rates[0]
    n_layers = int(wn_cfg.get("n_layers", 12))
    k = int(wn_cfg.get("kernel_size", 3))
    step = int(wn_cfg.get("dilation_rate_step", 1))
    max_log2 = wn_cfg.get("max_log2_dilation_rate", None)
    if k % 2 != 1 or any(int(n_channels * cf) % 2 for cf in chan_factors):
        raise AssertionError("WaveNetAE needs odd kernel_size and even n_channels")
    gate = wn_cfg.get("activation", "gtu")
    if gate not in _GATES:
        raise RuntimeError(f"WaveNetAE::error::unsupported wavenet activation {gate} selected. "
                           f"For gated units please select one of gtu, gfu, gsu, or glu.")
    if wn_cfg.get("n_ch_groups", 1) != 1 or wn_cfg.get("pre_cond_layer_channels") or \
            wn_cfg.get("disable_conditioning") or wn_cfg.get("use_equalized_lr"):
        raise NotImplementedError("WaveNet channel groups / pre-cond layers / equalized lr are not built yet")
    if wn_cfg.get("n_out_channels") is None:
        raise RuntimeError("WaveNetAE::error::n_out_channels parameter is required")
    dil = [2 ** ((i // step) % max_log2) if max_log2 is not None else 2 ** (i // step) for i in range(n_layers)]
    sigma = mc.get("pp_mod_subnet_noise_channel_sigma", 0.5)
    steps_per_frame = int(round(wn_rate / spect_rate))
    subharm = int(mc["wavetable_config"].get("add_subharm_chans", 0) or 0)             # tf_wavetable.py:212, :554-559
    pulse_pqmf_cfg = None
    if mc.get("pulse_channels_use_pqmf"):                                               # custom_pulsed_generator.py:499-501, :895
        pulse_pqmf_cfg = copy.deepcopy(mc.get("pulse_channels_multi_band_config") or {})
        if int(pulse_pqmf_cfg.get("subbands", 0)) != pch:
            raise RuntimeError(f"MBExWN::config_error::pulse_channels_multi_band_config.subbands "
                               f"{pulse_pqmf_cfg.get('subbands')} != pulse_channels {pch}")
        if int(pulse_pqmf_cfg["taps"]) % 2:
            raise AssertionError("The number of taps mush be even number.")             # tf_preprocess.py:44
    blocks = []
    for iwn, (u, cf, rate) in enumerate(zip(ups_factors, chan_factors, block_rates)):
        blocks.append(WaveNetSpec(
            name=f"PP_waveNetBlock_ups{mc['pp_mod_subnet_upsampling_factors'][iwn]}_{iwn}", c=int(n_channels * cf),
            c_in=(pch * (1 + subharm) + (1 if sigma else 0)) if iwn == 0 else int(wn_cfg["n_out_channels"]),
            c_out=int(wn_cfg["n_out_channels"]), n_layers=n_layers, k=k, dilations=list(dil), gate=_GATES[gate], cond_k=cond_k,
            cond_conv_up=int(rate // (spect_rate * cond_lin)), cond_lin_up=cond_lin,
            steps_per_frame=int(round(rate / spect_rate)), cond_cin=n_mel, causal=bool(mc.get("force_causal", False)), up=u))
    wn = blocks[0]

    win, fft = dsp_init.stft_sizes(sr, hop, mc.get("internal_win_size_s"), int(mc.get("internal_fft_over", 0)))
    fdb = mc.get("filter_max_db_range")
    plan = ModelPlan(
        sample_rate=sr, hop=hop, mel_channels=n_mel, pulse_rate_factor=prf, pulse_rate=pulse_rate,
        pulse_channels=pch, subbands=S, pulse_per_frame=pulse_per_frame, steps_per_frame=steps_per_frame,
        f0_min=float(mc.get("pp_min_frequency", 40.0)), f0_max=float(mc.get("pp_max_frequency", 600.0)),
        noise_sigma=float(sigma or 0.0), alpha=float(mc.get("alpha", 0.2)), use_prelu=use_prelu,
        pp_ops=pp_ops, ps_ops=ps_ops, n_ceps=n_ceps, wavenet=wn, post_name="MBExWNGen_PaNMPulseWaveNet_Post",
        pqmf_cfg=mb, stft_win=win, fft_size=fft,
        filter_max_log_range=(fdb / (20 * np.log10(np.exp(1)))) if fdb is not None else None,
        env_order_scale=mc.get("ps_env_order_scale"), wavetable_cfg=copy.deepcopy(mc["wavetable_config"]), norm=norm,
        ps_mode=ps_mode, ps_preserve_energy=preserve_energy, subharm=subharm, pulse_pqmf_cfg=pulse_pqmf_cfg,
        wavenet_blocks=blocks, sub_per_frame=sub_per_frame)
    # guard frames: the widest single tap of any block (guard rows are re-zeroed by every layer, so reaches do not add up);
    # the up-sampling conv between blocks reaches k - 1 = 2 rows at most
    half_span = max(d * (k - 1) // (1 if wn.causal else 2) for d in dil)
    plan.max_halo_frames = max(1, max(-(-max(half_span, 2 if b.up > 1 else 0) // b.steps_per_frame) for b in blocks))
    if finalize:
        finalize_plan(plan)
    return plan


def finalize_plan(plan: ModelPlan) -> ModelPlan:
    """Compute the init-time DSP constants (wavetables, PQMF bank, windows, lifters)."""
    plan.wavetables = dsp_init.build_wavetables(sample_rate=plan.pulse_rate, **plan.wavetable_cfg)
    mb = plan.pqmf_cfg
    if mb.get("max_band"):
        raise NotImplementedError("multi_band_config.max_band is not built yet")
    _, syn = dsp_init.pqmf_filters(mb["subbands"], mb["taps"], mb["cutoff_ratio"], mb["beta"])
    plan.pqmf_syn = syn
    plan.pqmf_poly, plan.pqmf_q, plan.pqmf_back = dsp_init.pqmf_polyphase(syn, mb["subbands"], mb["taps"])
    plan.window = dsp_init.hann_periodic(plan.stft_win)
    plan.inv_window = dsp_init.inverse_stft_window(plan.stft_win, plan.hop)
    plan.f0_smooth = dsp_init.f0_smoothing_kernel(plan.hop)
    if plan.env_order_scale:
        plan.lifter_log10f0, plan.lifters = dsp_init.cepstral_lifters(
            plan.env_order_scale, plan.sample_rate, plan.f0_min, plan.f0_max, plan.n_ceps)
    if plan.pulse_pqmf_cfg is not None:
        pq = plan.pulse_pqmf_cfg
        plan.pulse_pqmf_ana, _ = dsp_init.pqmf_filters(pq["subbands"], pq["taps"], pq["cutoff_ratio"], pq["beta"])
    if plan.norm is not None:
        nm = plan.norm
        hann = dsp_init.cosine_window("hann", nm.win).astype(np.float32)
        if nm.use_pinv:                                                                 # wavegen_1d.py:602-608
            basis = dsp_init.mel_filter_bank(nm.sample_rate, nm.fft_size, nm.n_mel, nm.fmin, nm.fmax)
            nm.proj = np.ascontiguousarray(np.linalg.pinv(basis).T.astype(np.float32))
            nm.proj_cols = int(nm.proj.shape[1])
            nm.proj_scale = float(1.0 / np.sqrt(np.sum(hann ** 2)))
        else:                                                                           # wavegen_1d.py:609-612
            mel_f = dsp_init.mel_frequencies(nm.n_mel + 2, nm.fmin, nm.fmax)
            nm.proj = ((mel_f[2:nm.n_mel + 2] - mel_f[:nm.n_mel]) / 2.).astype(np.float32)
            nm.proj_cols, nm.proj_scale = 0, 1.0
        sw = dsp_init.cosine_window("hann", nm.smooth_win).astype(np.float32)
        nm.smooth_window = sw ** 2 if nm.squared_win else sw
        nm.gwin = hann / np.sum(hann)
    return plan
