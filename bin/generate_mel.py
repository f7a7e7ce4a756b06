#!/usr/bin/env python3
"""Mel analysis of sound files on a B200 -- same command line and output as the reference's bin/generate_mel.py:28-94:
one pickled dictionary ``<name>.mell`` per input with the log-mel under ``mell`` (n_mels, frames) and the analysis
parameters the inverter's ``scale_mel`` needs.  All files are analysed as one ragged batch by the fused STFT/mel kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.exists(os.path.join(ROOT, "mbexwn_vocoder_b200")):
    sys.path.insert(0, ROOT)

from mbexwn_vocoder_b200 import fileio, get_config_file, list_models   # noqa: E402
from mbexwn_vocoder_b200 import config as cutils                       # noqa: E402


def main(input_audio_files, output_dir, model_id="VOICE", device=0):
    from mbexwn_vocoder_b200.analysis import MelAnalyzer, resample
    config_file = get_config_file(model_id_or_path=model_id)
    if not os.path.exists(config_file):
        raise FileNotFoundError(f"error::loading config file from {config_file}")
    hparams = cutils.read_config(config_file=config_file)
    pc = hparams['preprocess_config']
    if output_dir and not os.path.exists(output_dir):
        os.makedirs(output_dir)
    data_dict = {'nfft': pc["fft_size"], 'hoplen': pc["hop_size"], 'winlen': pc["win_size"], 'nmels': pc["mel_channels"],
                 'sr': pc['sample_rate'], 'fmin': pc['fmin'], 'fmax': pc['fmax'], 'lin_spec_offset': pc['lin_amp_off'],
                 'lin_spec_scale': pc['lin_amp_scale'], 'log_spec_offset': 0., 'log_spec_scale': pc['mel_amp_scale'],
                 "time_axis": 1}
    sounds = []
    for audio_file in input_audio_files:
        print(f"process {audio_file}", file=sys.stderr)
        snd, sr = fileio.read_audio(audio_file, dtype=np.float32)
        if snd.ndim > 1:
            # compute_mel_spectrogram_internal keeps the first row of its (batch, time) result (generate_mel.py:63)
            snd = snd[:, 0]
        if sr != pc['sample_rate']:
            snd = resample(snd, sr, pc['sample_rate'], axis=0)
        sounds.append(snd)
    mels = MelAnalyzer(pc, device=device)(sounds, do_post=False)
    outfiles = []
    for audio_file, mel in zip(input_audio_files, mels):
        out = dict(data_dict)
        out['mell'] = mel.T
        outfiles.append(os.path.join(output_dir, os.path.splitext(os.path.basename(audio_file))[0] + ".mell"))
        fileio.save_var(outfiles[-1], out)
    return outfiles


def build_parser():
    from argparse import ArgumentParser
    parser = ArgumentParser(description="create mel analysis from sound files using the configuration of a model")
    parser.add_argument("input_audio_files", nargs="*", help="input files to process")
    parser.add_argument("-o", "--output_dir", help="output directory where the .mell files will be stored")
    parser.add_argument("--model_id", default="VOICE", nargs="?", const="",
                        help="model identifier that is used to read the config file. If you do not specify an argument "
                             "after the --model_id flag the script will list all available models.")
    parser.add_argument("--device", default=0, type=int, help="CUDA device index (Def: %(default)s)")
    return parser


def cli(argv=None):
    args = build_parser().parse_args(argv)
    if not args.model_id:
        for kk, ll in list_models().items():
            for md in ll:
                print(f" - {kk}/{md}")
        return 0
    if not args.input_audio_files or not args.output_dir:
        print("generate_mel::error::input files and -o/--output_dir are required", file=sys.stderr)
        return 2
    main(**vars(args))
    return 0


if __name__ == "__main__":
    sys.exit(cli())
