"""B200-native MBExWN mel-inversion forward pass behind the reference's MELInverter API.

Mirrors MBExWN_NVoc/__init__.py:14-65 (model registry and config lookup).  The compute path lives in
``csrc/`` (hand-written sm_100a CUDA behind the C-ABI of ``include/mbexwn.h``); this package is the
host-side mirror of the reference's operator interface for that one path.
"""
import copy
import os
from typing import Union

mbexwn_version = (1, 2, 3)
b200_version = (0, 1, 0)

# Same registry as the reference (MBExWN_NVoc/__init__.py:21-31): domain -> ordered model names, the first
# entry of a domain is its default.  The released config.yaml/weights are a separate download that is not
# available offline, so each model directory here holds a *synthetic* config of the same architecture
# family and weights are random-initialised unless a weight file is present.
_mel_inv_models = {
    "SING": [
        "MBExWN_SIIConv_V71g_SING_IMP0_IMPORTmod_MCFG0_WNCHA320_DCHA32_1024_DPTACT0_ADLW0.1_GMCFG5_24kHz",
    ],
    "SPEECH": [
        "MBExWN_SIIConv_V71g_SPEECH_IMP0_IMPORTmod_MCFG0_WNCHA320_DCHA32_1024_DPTACT0_ADLW0.1_GMCFG5_24kHz",
    ],
    "VOICE": [
        "MBExWN_SIIConv_V71g_VOICE2_WNCHA340_IMP0_WNCHA340_IMPORTmod_MCFG0_WNCHA340_DCHA32_1024_DPTACT0_ADLW0.1_GMCFG0_24kHz",
    ],
}

# README.md:177-179 aliases of the reference
_aliases = {"MW-SI-FD": "SING", "MW-SP-FD": "SPEECH", "MW-VO-FD": "VOICE"}


def list_models(voice_type: Union[str, None] = None):
    """All mel-inverter models per voice class (MBExWN_NVoc/__init__.py:33-44; the argument is ignored there too)."""
    return copy.deepcopy(_mel_inv_models)


def get_config_file(model_id_or_path, verbose=False):
    """Resolve a model id or directory to its config.yaml (MBExWN_NVoc/__init__.py:47-65).

    An existing path is taken as the model directory; otherwise the id is matched as a substring of
    "DOMAIN/name".  Divergence from the reference: the first match in registry order
    wins (as README.md:174-177 documents) and an unknown id raises FileNotFoundError instead of
    UnboundLocalError.
    """
    from pathlib import Path

    model_dir = None
    if os.path.exists(model_id_or_path):
        model_dir = model_id_or_path
    else:
        key = _aliases.get(model_id_or_path, model_id_or_path)
        for domain, names in _mel_inv_models.items():
            for md in names:
                if model_dir is None and key in f"{domain}/{md}":
                    model_dir = Path(__file__).absolute().parent / "models" / md
    if model_dir is None:
        raise FileNotFoundError(f"error::no model matches id {model_id_or_path!r}")
    config_file = os.path.join(model_dir, "config.yaml")
    if not os.path.exists(config_file):
        raise FileNotFoundError(f"error::loading config file from {config_file}")
    return config_file
