#!/usr/bin/env python
"""Convert model weights between the reference's TensorFlow checkpoint (weights.tf.index / .data-*) and this package's
weights.npz, without TensorFlow (mbexwn_vocoder_b200/tf_checkpoint.py).

    python tools/convert_checkpoint.py tf2npz <model_dir>      # weights.tf -> weights.npz beside config.yaml
    python tools/convert_checkpoint.py npz2tf <model_dir>      # weights.npz (or the synthetic init) -> weights.tf
    python tools/convert_checkpoint.py list   <model_dir>      # keys, dtypes and shapes of weights.tf
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbexwn_vocoder_b200 import get_config_file, tf_checkpoint as T, weights as W      # noqa: E402
from mbexwn_vocoder_b200.config import read_config                                      # noqa: E402
from mbexwn_vocoder_b200.plan import build_plan                                         # noqa: E402


def main(argv):
    if len(argv) != 3 or argv[1] not in ("tf2npz", "npz2tf", "list"):
        print(__doc__, file=sys.stderr)
        return 2
    config_file = get_config_file(argv[2])
    model_dir = os.path.dirname(config_file)
    hp = read_config(config_file)
    prefix = os.path.join(model_dir, "weights.tf")
    if argv[1] == "list":
        r = T.BundleReader(prefix)
        for k in sorted(r.keys()):
            e = r.entries[k]
            print(f"{k}  dtype={e.dtype} shape={e.shape} bytes={e.size}")
        return 0
    plan = build_plan(hp, finalize=False)
    if argv[1] == "tf2npz":
        W.save(os.path.join(model_dir, "weights.npz"), T.import_weights(prefix, plan))
        print(f"wrote {os.path.join(model_dir, 'weights.npz')}")
    else:
        npz = os.path.join(model_dir, "weights.npz")
        w = W.load(npz) if os.path.exists(npz) else W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))
        W.check(plan, w)
        T.export_weights(prefix, hp, w)
        print(f"wrote {prefix}.index / .data-00000-of-00001")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
