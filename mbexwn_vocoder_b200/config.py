"""YAML model-config reader (restates MBExWN_NVoc/vocoder/model/config_utils.py:33-60, :271-339).

Only the load path is here: environment / home expansion, ``<@CONFIG_DIR@/file:key:...>`` includes and
``__defaults__`` completion.  The CLI override mini-language and the training validators of the reference
(config_utils.py:102-229, :353-515) are out of scope.
"""
from __future__ import annotations

import io
import os
import re
from copy import deepcopy

import numpy as np
import yaml

# the reference maps dtype strings to tf/np dtypes; without TensorFlow the tf entries map to NumPy dtypes
_type_map = {
    "tf.float32": np.float32,
    "tf.float16": np.float16,
    "np.float32": np.float32,
    "np.float16": np.float16,
    "None": None,
}

_include_pat = re.compile(r"<@CONFIG_DIR@/(.*)>$")


def _expand(value, base_dir):
    if isinstance(value, str):
        if value in _type_map:
            return _type_map[value]
        if "$" in value:
            value = os.path.expandvars(value)
        if "~" in value:
            value = os.path.expanduser(value)
        stripped = value.strip()
        mapped = _include_pat.sub(lambda m: f"{base_dir}/{m.group(1)}", stripped)
        if mapped != stripped:
            file_name, *keys = mapped.split(":")
            value = read_config(file_name, config_base_dir=base_dir)
            for key in keys:
                value = value[key]
        return value
    if isinstance(value, dict):
        for key, sub in value.items():
            value[key] = _expand(sub, base_dir)
    elif isinstance(value, list):
        for i in range(len(value)):
            value[i] = _expand(value[i], base_dir)
    return value


def _fill_defaults(config):
    """In-place ``__defaults__`` completion for dicts and lists of dicts (config_utils.py:271-312)."""
    snapshot = deepcopy(config)
    for key, val in snapshot.items():
        if key == "__defaults__":
            for dk, dv in val.items():
                if dk not in config:
                    config[dk] = dv
            config.pop("__defaults__")
        elif isinstance(val, dict):
            _fill_defaults(config[key])
        elif isinstance(val, list):
            defaults, where = None, None
            for i, entry in enumerate(val):
                if isinstance(entry, dict) and len(entry) == 1 and "__defaults__" in entry:
                    if where is not None:
                        raise RuntimeError(f"read_config::error::multiple __defaults__ entries in list {val}")
                    defaults, where = deepcopy(entry["__defaults__"]), i
            if where is not None:
                del config[key][where]
                for entry in config[key]:
                    if not isinstance(entry, dict):
                        raise RuntimeError(f"read_config::error::cannot use default values from {defaults} "
                                           f"for list entries that are not dicts {entry}")
                    for dk, dv in defaults.items():
                        if dk not in entry:
                            entry[dk] = dv
            for entry in config[key]:
                if isinstance(entry, dict):
                    _fill_defaults(entry)


def read_config(config_file, config_base_dir=None):
    """Read one YAML file (or the concatenation of several) into a dict (config_utils.py:314-339)."""
    if config_base_dir is None:
        config_base_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
    files = config_file if isinstance(config_file, (list, tuple)) else [config_file]
    stream = io.StringIO()
    for name in files:
        with open(name, "r") as fi:
            stream.write(fi.read())
    stream.seek(0)
    config = yaml.safe_load(stream)
    for key, val in config.items():
        config[key] = _expand(val, config_base_dir)
    _fill_defaults(config)
    return config
