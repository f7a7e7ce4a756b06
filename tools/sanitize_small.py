#!/usr/bin/env python
"""Tiny end-to-end runs for compute-sanitizer (memcheck / racecheck): every precision, a ragged batch, the analysis kernel.
`sanitize_small.py blocks` runs only the multi-block / up-sampling WaveNet variants (two model directories written to /tmp)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbexwn_vocoder_b200.mel_inverter import MELInverter
from oracle.forward import synthetic_mel
from oracle.analysis import synthetic_audio

mels = [synthetic_mel(t, i) for i, t in enumerate((9, 1, 14))]


def multi_block_variants():
    import tempfile
    import yaml
    from mbexwn_vocoder_b200 import get_config_file
    variants = {"blocks_2x1": {"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [2, 1],
                               "pp_mod_subnet_channel_factors": [0.5, 0.25]},
                "blocks_1x2": {"pulse_channels": 10, "pp_mod_subnet_upsampling_factors": [1, 2],
                               "pp_mod_subnet_channel_factors": [0.25, 0.5]},
                "lifter_causal": {"ps_env_order_scale": 1.5, "force_causal": True}}       # F0-dependent cepstral lifter
    for name, extra in variants.items():
        cfg = yaml.safe_load(open(get_config_file("SPEECH")))
        cfg["mbexwn_config"].update(extra)
        d = tempfile.mkdtemp(prefix=name)
        yaml.safe_dump(cfg, open(os.path.join(d, "config.yaml"), "w"))
        mb = MELInverter(d, device=0, precision="f16f8", allow_synthetic_weights=True)
        for prec in ("f16f8", "bf16", "fp32"):
            mb.precision = prec
            out = mb.synth_batch(mels)
            assert all(np.isfinite(w).all() for w in out), (name, prec)
            print(name, prec, "ok", [w.shape for w in out])


if len(sys.argv) > 1 and sys.argv[1] == "blocks":
    multi_block_variants()
    sys.exit(0)
inv = MELInverter("SPEECH", device=0, precision="f16f8", allow_synthetic_weights=True)
for prec in ("f16f8", "bf16x3", "bf16", "fp32"):
    inv.precision = prec
    out = inv.synth_batch(mels)
    assert all(np.isfinite(w).all() for w in out), prec
    print(prec, "ok", [w.shape for w in out])
dd = inv.generate_mel_from_snd(synthetic_audio(2000, 0), 24000)
print("analysis ok", dd["mell"].shape)
long = inv.synth_long_from_mel(synthetic_mel(45, 3), chunk_frames=20)
print("long ok", long.shape)
multi_block_variants()
