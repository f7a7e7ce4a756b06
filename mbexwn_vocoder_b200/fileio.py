"""File formats either side of the hot path: the reference's pickled ``.mell`` analysis dictionaries
(MBExWN_NVoc/fileio/iovar.py:36-106, written by bin/generate_mel.py:41-74) and sound files.

The reference reads / writes audio with pysndfile (libsndfile), which is not available here; WAV (PCM 16 / 24 / 32 bit and
IEEE float) is handled with NumPy only and every other container raises with a clear message.
"""
from __future__ import annotations

import gzip
import pickle
import struct
from typing import Tuple

import numpy as np


def save_var(filename: str, data, protocol: int = -1) -> None:
    """Pickle ``data`` to ``filename`` (gzip when it ends in .gz) -- iovar.save_var without the dill option."""
    opener = gzip.open if filename.endswith(".gz") else open
    with opener(filename, "wb") as f:
        pickle.dump(data, f, protocol)


def load_var(filename: str):
    """iovar.load_var: un-pickle, retrying with latin1 for files written under Python 2 (iovar.py:90-96)."""
    opener = gzip.open if filename.endswith(".gz") else open
    try:
        with opener(filename, "rb") as f:
            return pickle.load(f)
    except UnicodeDecodeError:
        with opener(filename, "rb") as f:
            return pickle.load(f, encoding="latin1")


_WAVE_FORMAT_PCM, _WAVE_FORMAT_FLOAT, _WAVE_FORMAT_EXTENSIBLE = 1, 3, 0xFFFE


def read_audio(path: str, dtype=np.float32) -> Tuple[np.ndarray, int]:
    """(samples [, channels]) in [-1, 1) and the sample rate of a RIFF/WAVE file."""
    data = open(path, "rb").read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise RuntimeError(f"{path}: only RIFF/WAVE files can be read here (the reference uses libsndfile, which is "
                           f"not installed)")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == _WAVE_FORMAT_EXTENSIBLE and len(body) >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise RuntimeError(f"{path}: malformed WAVE file (fmt or data chunk missing)")
    tag, ch, sr, bits = fmt
    if tag == _WAVE_FORMAT_FLOAT and bits in (32, 64):
        x = np.frombuffer(payload, dtype="<f4" if bits == 32 else "<f8").astype(dtype)
    elif tag == _WAVE_FORMAT_PCM and bits == 16:
        x = np.frombuffer(payload, dtype="<i2").astype(dtype) / 32768.0
    elif tag == _WAVE_FORMAT_PCM and bits == 32:
        x = np.frombuffer(payload, dtype="<i4").astype(np.float64) / 2147483648.0
    elif tag == _WAVE_FORMAT_PCM and bits == 24:
        raw = np.frombuffer(payload[:len(payload) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
        x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / 8388608.0
    elif tag == _WAVE_FORMAT_PCM and bits == 8:
        x = (np.frombuffer(payload, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
    else:
        raise RuntimeError(f"{path}: unsupported WAVE encoding (format tag {tag}, {bits} bits)")
    x = x.astype(dtype, copy=False)
    if ch > 1:
        x = x[:x.size // ch * ch].reshape(-1, ch)
    return x, int(sr)


def write_audio(path: str, data: np.ndarray, rate: int, format: str = "wav", enc: str = "float32") -> None:
    """Write mono / multi-channel float samples; ``enc`` is pcm16 or float32.  Only the wav container is built in."""
    if format.lower() not in ("wav", "wave"):
        raise RuntimeError(f"cannot write '{format}' files: the reference relies on libsndfile (pysndfile) for that, "
                           f"which is not installed; use --format wav")
    x = np.asarray(data)
    ch = 1 if x.ndim == 1 else x.shape[1]
    if enc == "pcm16":
        body = np.round(np.clip(x, -1.0, 32767.0 / 32768.0) * 32768.0).astype("<i2").tobytes()
        tag, bits = _WAVE_FORMAT_PCM, 16
    elif enc == "float32":
        body = x.astype("<f4").tobytes()
        tag, bits = _WAVE_FORMAT_FLOAT, 32
    else:
        raise RuntimeError(f"unsupported encoding {enc}")
    fmt = struct.pack("<HHIIHH", tag, ch, int(rate), int(rate) * ch * bits // 8, ch * bits // 8, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if tag == _WAVE_FORMAT_FLOAT:
        chunks += b"fact" + struct.pack("<II", 4, x.shape[0])
    chunks += b"data" + struct.pack("<I", len(body)) + body + (b"\0" if len(body) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
