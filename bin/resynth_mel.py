#!/usr/bin/env python3
"""Mel inversion from pickled mel files on a B200 -- same command line as the reference's bin/resynth_mel.py:34-135.

    resynth_mel.py [model_id] -i a.mell b.mell ... -o out_dir [--format wav] [-v] [-q]

Differences from the reference, all on the host side of the boundary: every input file is scaled with
``MELInverter.scale_mel`` and the whole set is synthesised as ONE variable-length batch (the reference loops file by file);
``-g`` / ``-nt`` are accepted and ignored (the path always runs on the GPU); audio is written as WAV unless a libsndfile
binding is importable (``--format flac`` is the reference's default, kept when ``soundfile`` exists).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.exists(os.path.join(ROOT, "mbexwn_vocoder_b200")):
    sys.path.insert(0, ROOT)

from mbexwn_vocoder_b200 import fileio, list_models, mel_inverter      # noqa: E402


def default_format() -> str:
    try:
        import soundfile  # noqa: F401
        return "flac"
    except ImportError:
        return "wav"


def _write(outfile, audio, rate, fmt):
    if fmt.lower() in ("wav", "wave"):
        fileio.write_audio(outfile, audio, rate, "wav")
        return
    try:
        import soundfile
    except ImportError:
        fileio.write_audio(outfile, audio, rate, fmt)               # raises with the explanation
        return
    soundfile.write(outfile, audio, rate, format=fmt.upper())


def main(model_id, input_mell_files, output_dir, use_gpu=True, format=None, verbose=False, seed=42, num_threads=2,
         quiet=False, precision="f16f8", device=0, devices=None, max_batch_frames=32768, synthetic_weights=False):
    format = format or default_format()
    if seed >= 0:
        np.random.seed(seed)
    MelInv = mel_inverter.MELInverter(model_id_or_path=model_id, device=device, precision=precision,
                                      seed=seed if seed >= 0 else 42, verbose=verbose, devices=devices,
                                      allow_synthetic_weights=True if synthetic_weights else None)
    if output_dir and not os.path.exists(output_dir):
        os.makedirs(output_dir)

    mels, outfiles = [], []
    for mell_file in input_mell_files:
        outfile = os.path.join(output_dir, "syn_" + os.path.splitext(os.path.basename(mell_file))[0] + "." + format)
        if verbose:
            print(f"load mell  from {mell_file}", file=sys.stderr)
        dd = fileio.load_var(mell_file)
        mels.append(MelInv.scale_mel(dd, verbose=verbose)[0])
        outfiles.append(outfile)

    start_time = time.time()
    # the reference loops over the files one by one (bin/resynth_mel.py:72-104); here the set is LPT-sharded over the devices
    # and cut into batches under a frame budget, so memory does not grow with the number of files
    audios = MelInv.synth_many(mels, max_batch_frames=max_batch_frames)
    end_time = time.time()
    total = sum(a.size for a in audios)
    if verbose:
        print(f"    synthesized {len(audios)} file(s), {total} samples in {end_time - start_time:.3f}s "
              f"({total / (end_time - start_time):.2f}Hz)", file=sys.stderr)

    for mell_file, outfile, log_mel, syn_audio in zip(input_mell_files, outfiles, mels, audios):
        if not quiet:
            print(f"synthesize {mell_file} into {outfile}", file=sys.stderr)
        if verbose:
            mel_resyn = MelInv.generate_mel_from_snd(syn_audio, srate=MelInv.srate)['mell'].T
            mell_err = mel_inverter.log_to_db * np.mean(np.abs(log_mel - mel_resyn[:log_mel.shape[0]]))
            print(f"    {syn_audio.size} samples, mel_error: {mell_err:.3f}dB", file=sys.stderr)
        if np.max(np.abs(syn_audio)) > 1:
            norm = 0.99 / np.max(np.abs(syn_audio))
            print(f'    to prevent clipping you would need to normalize {outfile} by {norm:.3f}', file=sys.stderr)
        if verbose:
            print(f"    save audio under {outfile}", file=sys.stderr)
        _write(outfile, syn_audio, MelInv.srate, format)
    return outfiles


def build_parser():
    from argparse import ArgumentParser
    parser = ArgumentParser(description="pass mel spectrograms through the MBExWN mel inverter on a B200")
    parser.add_argument("model_id", default=None, nargs="?", const=None,
                        help="model identifier. If not given the script will list all known model names. You don't need "
                             "the full model name: the first model containing the given identifier is used. A path to a "
                             "model directory (config.yaml + weights) is accepted as well.")
    parser.add_argument("-i", "--input_mell_files", nargs="+", help="list of mell spectra stored in pickle files")
    parser.add_argument("-o", "--output_dir", help="output directory where synthetic sounds will be stored")
    parser.add_argument("--format", default=None, help="file format for generated audio files (Def: flac when a "
                                                       "libsndfile binding is importable, else wav)")
    parser.add_argument("-nt", "--num_threads", default=2, type=int, help="accepted for compatibility (ignored)")
    parser.add_argument("-g", "--use_gpu", action="store_true", help="accepted for compatibility (always on)")
    parser.add_argument("-v", "--verbose", action="store_true", help="display verbose progress info")
    parser.add_argument("-q", "--quiet", action="store_true", help="dont display progress")
    parser.add_argument("--precision", default="f16f8", choices=["f16f8", "bf16x3", "bf16", "fp32"],
                        help="arithmetic of the WaveNet contractions (Def: %(default)s, fp32-accurate)")
    parser.add_argument("--device", default=0, type=int, help="CUDA device index (Def: %(default)s)")
    parser.add_argument("--devices", default=None, type=int, nargs="+",
                        help="several CUDA devices: the files are sharded over them by length (longest-processing-time first)")
    parser.add_argument("--max_batch_frames", default=32768, type=int,
                        help="mel frames (incl. guard frames) per forward call and GPU (Def: %(default)s, about 7 minutes of audio)")
    parser.add_argument("--synthetic_weights", action="store_true",
                        help="accept a model directory without weights and run on random-initialised ones (tests only)")
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    if not args.model_id:
        print("Please select one of the following models for mel inversion.\nYou don't need to select with a full ID. "
              "The first model containing the model_id you provide will be selected.\nFor example just specifying SPEECH "
              "will select the default SPEECH model.")
        for kk, ll in list_models().items():
            for md in ll:
                print(f" - {kk}/{md}")
        return 0
    if not args.input_mell_files or not args.output_dir:
        print("resynth_mel::error::-i/--input_mell_files and -o/--output_dir are required", file=sys.stderr)
        return 2
    main(**vars(args))
    return 0


if __name__ == "__main__":
    sys.exit(cli())
