#!/usr/bin/env python
"""Install the reference's released models into this package (the counterpart of
scripts/download_and_install_MBExWN_pretrained_models.sh, which needs network access).

    python tools/install_models.py SIIConv_pretrained_MBExWN_models.zip      # the zip of the reference's download script
    python tools/install_models.py /path/to/MBExWN_NVoc/models               # or an unpacked models directory

The zip holds ./MBExWN_NVoc/models/<model name>/{config.yaml, weights.tf.index, weights.tf.data-*}
(scripts/create_SIIConv_pretrained_models_zip.sh:5).  Every model directory is copied to mbexwn_vocoder_b200/models/<name>/,
replacing the synthetic config of the same name; the checkpoint is then read without TensorFlow (tf_checkpoint.py) and
checked against the plan of its config.  --convert additionally writes weights.npz beside it.
"""
import argparse
import os
import shutil
import sys
import tempfile
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_model_dirs(top):
    out = []
    for dirpath, _, files in os.walk(top):
        if "config.yaml" in files:
            out.append(dirpath)
    return sorted(out)


def install(src, dest_root, convert=False, check=True):
    from mbexwn_vocoder_b200 import tf_checkpoint as T, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    installed = []
    for d in find_model_dirs(src):
        name = os.path.basename(os.path.normpath(d))
        dest = os.path.join(dest_root, name)
        os.makedirs(dest, exist_ok=True)
        for f in os.listdir(d):
            if f == "config.yaml" or f.startswith("weights."):
                shutil.copy2(os.path.join(d, f), os.path.join(dest, f))
        msg = "copied"
        if check or convert:
            hp = read_config(os.path.join(dest, "config.yaml"))
            plan = build_plan(hp, finalize=False)
            prefix = os.path.join(dest, "weights.tf")
            if os.path.exists(prefix + ".index"):
                w = T.import_weights(prefix, plan)
                msg = f"checkpoint ok ({len(w)} tensors)"
                if convert:
                    W.save(os.path.join(dest, "weights.npz"), w)
                    msg += ", weights.npz written"
            elif os.path.exists(os.path.join(dest, "weights.npz")):
                W.check(plan, W.load(os.path.join(dest, "weights.npz")))
                msg = "weights.npz ok"
            else:
                msg = "no weights found (synthetic initialisation will be used)"
        print(f"{name}: {msg}")
        installed.append(dest)
    if not installed:
        raise FileNotFoundError(f"no model directory (config.yaml) found under {src}")
    return installed


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("source", help="zip file of the reference's download script, or a directory holding model directories")
    ap.add_argument("--dest", default=os.path.join(ROOT, "mbexwn_vocoder_b200", "models"))
    ap.add_argument("--convert", action="store_true", help="also write weights.npz")
    ap.add_argument("--no-check", action="store_true", help="copy only, do not parse the checkpoints")
    args = ap.parse_args(argv)
    if os.path.isdir(args.source):
        install(args.source, args.dest, args.convert, not args.no_check)
    else:
        with tempfile.TemporaryDirectory() as tmp:
            with zipfile.ZipFile(args.source) as z:
                z.extractall(tmp)
            install(tmp, args.dest, args.convert, not args.no_check)
    return 0


if __name__ == "__main__":
    sys.exit(main())
