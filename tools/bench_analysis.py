#!/usr/bin/env python
"""Device time of the analysis kernel (audio -> log-mel) on a config-2 sized batch: audio-s/s, achieved GB/s on the
algorithmic bytes (4 B per sample in, 4 * n_mel B per frame out) and the host-buffer (H2D + kernel + D2H) rate."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mbexwn_vocoder_b200 import get_config_file                    # noqa: E402
from mbexwn_vocoder_b200.analysis import MelAnalyzer               # noqa: E402
from mbexwn_vocoder_b200.config import read_config                 # noqa: E402


def main():
    batch, seconds = int(os.environ.get("B", 64)), float(os.environ.get("S", 5.0))
    pc = read_config(get_config_file("SPEECH"))["preprocess_config"]
    an = MelAnalyzer(pc, device=0)
    n = int(seconds * pc["sample_rate"])
    st = an.prepare([n] * batch)
    g = torch.Generator().manual_seed(0)
    audio = (0.1 * torch.randn(batch * n, generator=g)).pin_memory()
    mel = torch.empty(st["mel"].shape, dtype=torch.float32).pin_memory()
    st["audio"].copy_(audio)
    for _ in range(3):
        an.run(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        an.run(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    for _ in range(2):
        an.run_host(st, audio, mel)
    e0.record()
    for _ in range(10):
        an.run_host(st, audio, mel)
    e1.record()
    torch.cuda.synchronize()
    ms_host = e0.elapsed_time(e1) / 10
    frames = st["mel"].shape[0]
    alg_bytes = batch * n * 4 + frames * an.n_mel * 4
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    print(json.dumps({"kernel": "mel_analysis2048_kernel", "batch": batch, "seconds": seconds, "frames": frames,
                      "ms": ms, "audio_s_per_s": batch * seconds / (ms * 1e-3), "algorithmic_gbs": alg_bytes / ms / 1e6,
                      "frac_hbm": alg_bytes / ms / 1e6 / hbm, "fft_gflops": frames / 2 * 5 * 2048 * 11 / ms / 1e6,
                      "host_ms": ms_host, "host_audio_s_per_s": batch * seconds / (ms_host * 1e-3)}))


if __name__ == "__main__":
    main()
