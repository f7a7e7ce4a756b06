"""Caller-level pieces either side of the hot path: .mell pickles, WAV io, the two command lines."""
import importlib.util
import os
import pickle
import struct
import sys
import wave

import numpy as np
import pytest

from mbexwn_vocoder_b200 import fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_script(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "bin", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mell_pickle_round_trip(tmp_path):
    dd = {"mell": np.random.default_rng(0).standard_normal((80, 12)).astype(np.float32), "sr": 24000, "hoplen": 300,
          "nfft": 2048, "fmin": 0, "fmax": 12000, "time_axis": 1}
    for name in ("a.mell", "b.mell.gz"):
        path = str(tmp_path / name)
        fileio.save_var(path, dd)
        back = fileio.load_var(path)
        assert sorted(back) == sorted(dd) and np.array_equal(back["mell"], dd["mell"]) and back["sr"] == 24000
    # a plain pickle written by the reference's iovar.save_var (std pickle, protocol -1) loads as is
    with open(tmp_path / "ref.mell", "wb") as f:
        pickle.dump(dd, f, -1)
    assert np.array_equal(fileio.load_var(str(tmp_path / "ref.mell"))["mell"], dd["mell"])


def test_wav_round_trip_and_stdlib_compatibility(tmp_path):
    x = (0.5 * np.sin(2 * np.pi * 440 * np.arange(2400) / 24000)).astype(np.float32)
    p32 = str(tmp_path / "f.wav")
    fileio.write_audio(p32, x, 24000)
    y, sr = fileio.read_audio(p32)
    assert sr == 24000 and y.dtype == np.float32 and np.array_equal(y, x)
    p16 = str(tmp_path / "i.wav")
    fileio.write_audio(p16, x, 24000, enc="pcm16")
    y16, _ = fileio.read_audio(p16)
    assert np.abs(y16 - x).max() <= 1.0 / 32768
    with wave.open(p16, "rb") as w:                           # the stdlib reader agrees on the header
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 2400)
    # a file produced by the stdlib writer, stereo, is read back as (samples, channels)
    ps = str(tmp_path / "s.wav")
    with wave.open(ps, "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(struct.pack("<8h", 0, 16384, -16384, 32767, 1, -1, 100, -100))
    s, sr = fileio.read_audio(ps)
    assert sr == 16000 and s.shape == (4, 2) and s[0, 1] == 0.5 and s[1, 0] == -0.5
    with pytest.raises(RuntimeError, match="libsndfile"):
        fileio.write_audio(str(tmp_path / "x.flac"), x, 24000, format="flac")
    open(tmp_path / "junk.wav", "wb").write(b"not a wave file at all")
    with pytest.raises(RuntimeError, match="RIFF"):
        fileio.read_audio(str(tmp_path / "junk.wav"))


def test_command_lines_list_models_and_validate(capsys):
    rs, gm = _load_script("resynth_mel"), _load_script("generate_mel")
    assert rs.cli([]) == 0
    out = capsys.readouterr().out
    assert " - SPEECH/MBExWN_SIIConv_V71g_SPEECH" in out and " - VOICE/" in out and " - SING/" in out
    assert gm.cli(["--model_id"]) == 0
    assert " - VOICE/" in capsys.readouterr().out
    assert rs.cli(["SPEECH"]) == 2 and gm.cli(["x.wav"]) == 2
    args = rs.build_parser().parse_args(["SPEECH", "-i", "a.mell", "b.mell", "-o", "out", "-nt", "4", "-g", "-v"])
    assert args.model_id == "SPEECH" and args.input_mell_files == ["a.mell", "b.mell"] and args.num_threads == 4


@pytest.mark.gpu
def test_generate_then_resynth_end_to_end(tmp_path):
    """wav -> generate_mel -> .mell -> resynth_mel -> wav, through both command lines on the GPU."""
    from oracle.analysis import synthetic_audio
    rs, gm = _load_script("resynth_mel"), _load_script("generate_mel")
    wavs = []
    for i, n in enumerate((12000, 7000)):
        wavs.append(str(tmp_path / f"in{i}.wav"))
        fileio.write_audio(wavs[-1], synthetic_audio(n, i), 24000)
    mell_dir, out_dir = str(tmp_path / "mell"), str(tmp_path / "syn")
    assert gm.cli(wavs + ["-o", mell_dir, "--model_id", "SPEECH"]) == 0
    mells = sorted(os.path.join(mell_dir, f) for f in os.listdir(mell_dir))
    assert [os.path.basename(m) for m in mells] == ["in0.mell", "in1.mell"]
    dd = fileio.load_var(mells[0])
    assert dd["mell"].shape == (80, 41) and dd["sr"] == 24000 and dd["hoplen"] == 300 and dd["nfft"] == 2048
    assert rs.cli(["SPEECH", "-i"] + mells + ["-o", out_dir, "--format", "wav", "-v"]) == 0
    for name, frames in (("syn_in0.wav", 41), ("syn_in1.wav", 24)):
        y, sr = fileio.read_audio(os.path.join(out_dir, name))
        assert sr == 24000 and y.shape == (frames * 300,) and np.all(np.isfinite(y)) and np.abs(y).max() > 0
