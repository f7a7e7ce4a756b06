#!/usr/bin/env python
"""Pin the forward oracle against the REAL reference (needs TensorFlow >= 2.5 and the reference package importable;
neither exists in the build image, so this script is committed for a maintainer to run elsewhere).

    python tests/golden/make_tf_reference_goldens.py /path/to/MBExWN_Vocoder [SPEECH]

It (1) copies this package's synthetic config.yaml into a scratch model directory, (2) writes this package's synthetic
weights there as a TensorFlow checkpoint with the reference's object graph (tf_checkpoint.export_weights), (3) loads the
directory with the unmodified ``MBExWN_NVoc.mel_inverter.MELInverter``, (4) replaces the generator's ``tf.random.normal``
draw (custom_pulsed_generator.py:906) by the seeded noise of SURVEY.md 8d, and (5) stores mel, noise, F0, excitation and
waveform as ``tests/golden/tf_reference_<model>.npz``.  ``tests/test_oracle.py::test_oracle_against_tf_reference`` then
compares the restated forward with that file (skipped while the file is absent).
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main(argv):
    ref_root, model_id = argv[1], (argv[2] if len(argv) > 2 else "SPEECH")
    sys.path.insert(0, ref_root)
    import tensorflow as tf
    from MBExWN_NVoc import mel_inverter as ref_mi
    from MBExWN_NVoc.vocoder.model import custom_pulsed_generator as ref_gen

    from mbexwn_vocoder_b200 import get_config_file, tf_checkpoint as T, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import synthetic_mel, synthetic_noise

    cfg = get_config_file(model_id)
    hp = read_config(cfg)
    plan = build_plan(hp, finalize=False)
    w = W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))
    frames = 48
    mel = synthetic_mel(frames, 0)[None]
    noise = synthetic_noise(frames * plan.steps_per_frame, 0)[None]           # (1, T*steps, 1)
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(cfg, os.path.join(d, "config.yaml"))
        T.export_weights(os.path.join(d, "weights.tf"), hp, w)
        inv = ref_mi.MELInverter(d)
        real_normal = tf.random.normal

        def fixed_normal(shape, *a, **k):
            if tuple(shape) == noise.shape:
                return tf.constant(noise)
            return real_normal(shape, *a, **k)
        ref_gen.tf.random.normal = fixed_normal
        try:
            signals, pp = inv.model.infer(tf.constant(mel), synth_length=frames * plan.hop, return_F0=True,
                                          return_components=True)
        finally:
            ref_gen.tf.random.normal = real_normal
    out = {"mel": mel, "noise": noise, "waveform": np.asarray(signals[0])}
    for name, value in pp:
        out[str(name)] = np.asarray(value)
    path = os.path.join(HERE, f"tf_reference_{model_id}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: " + ", ".join(f"{k}{v.shape}" for k, v in out.items()))


if __name__ == "__main__":
    main(sys.argv)
