#!/usr/bin/env python
"""Headline benchmark: audio-seconds generated per wall-second (24 kHz) of the MBExWN mel-inversion forward pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f16f8|bf16x3|bf16|fp32] [--workload config2|config3]
    python bench.py --impl reference ...        # the restated reference forward on the box's host cores

One "step" = one pass of the hot path over one batch of synthetic mels (BASELINE.json configs[1]: MW-SP-FD,
batch 64 x 5 s, fp32-accurate = f16f8 split-precision tensor-core path: fp16 product + two e4m3 correction products).  N > 1 (torchrun): every rank runs its own batch of the
same size (independent utterances, no data-path collective; weak scaling); time = max over ranks.

JSON keys beyond the base contract:
  value     whole-job audio-s/s with inputs already resident in HBM (device events around K steps)
  e2e       the same metric through the product's Python API, NumPy in -> NumPy out: MELInverter.synth_stream over the
            step's batches (layout scatter into pinned memory, H2D, forward, D2H, per-utterance copies out; two buffer sets,
            the host work of step i +- 1 under the kernels of step i); e2e.cabi = the C-ABI calls alone on pre-filled pinned
            buffers (mbexwn_forward_host_begin/_wait), e2e.serial = one blocking mbexwn_forward_host per step
  config4   (sub-record) BASELINE.json configs[3]: 8192 utterances of 1-30 s, LPT-sharded over the ranks, host prep and the
            final host gather inside the timing (strong scaling; multi_gpu.run_shard)
  roofline  the dominant kernel (wn_gemm_kernel<EPI_GATE>: dilated-conv tap-GEMM + gate epilogue, one launch per WaveNet
            layer): its algorithmic TFLOP/s over its average launch duration, from CUDA events recorded inside the library
            on the launch stream around every launch (a separate pass after the headline timing)
  stages    device ms per stage + achieved GB/s (algorithmic bytes) for the HBM-bound stages
  cpu_baseline  CPU oracle (restated reference forward, torch-CPU fp32) on a bounded sample of the workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def ncu_traffic(workload, precision, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the newest `ncu --set full`
    capture committed under profiles/ (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None without one."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        table = json.load(open(path))
    except Exception:
        return None
    hit = table.get(f"{workload}/{precision}/{kernel}")
    return None if hit is None else float(hit["dram_bytes_per_launch"])

WORKLOADS = {
    # name: (model id, batch, frames, description)
    "config2": ("SPEECH", 64, 400, "MW-SP-FD batch 64 x 5 s synthetic mels"),
    "config3": ("VOICE", 256, 800, "MW-VO-FD batch 256 x 10 s synthetic mels"),
    "config1": ("SPEECH", 1, 400, "MW-SP-FD batch 1 x 5 s synthetic mel"),
}
CONFIG4 = {"model": "SING", "utts": 8192, "min_frames": 80, "max_frames": 2400, "seed": 1,
           "desc": "MW-SI-FD 8192 synthetic utterances of mixed length 1-30 s"}


def config_dict(args, world, batch, frames, desc, extra=None):
    """`config` of the JSON line: identical for the own arm and the reference arm (the driver compares them)."""
    # residual stream of one WaveNet layer: 20 rows per frame x 4 * 320 bytes -- what every layer reads and re-writes
    stream_mb = batch * frames * 20 * 1280 / 1e6
    c = {"workload": f"{args.workload}: {desc}", "per_gpu_batch": batch, "frames": frames, "precision": args.precision,
         "parallelism": f"dp{world} (independent utterances, no collective)",
         "l2": (f"no flush between steps: each WaveNet layer streams {stream_mb:.0f} MB in and out, "
                + ("several times the 126 MB L2" if stream_mb > 2 * 126 else "which fits the 126 MB L2 -- an L2-warm figure"))}
    if extra:
        c.update(extra)
    return c



def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def synthetic_batch(batch, frames, steps_per_frame, seed0=0):
    """Synthetic scaled log-mels of SURVEY.md 8d (clip(N(-4,2)), 5-frame box smoothing) and N(0,1) noise draws."""
    import torch
    mels, noise = [], []
    for u in range(batch):
        g = torch.Generator().manual_seed(1234 + seed0 + u)
        x = torch.clamp(torch.randn(frames + 4, 80, generator=g) * 2.0 - 4.0, min=float(np.log(1e-5)), max=2.0)
        mels.append(x.unfold(0, 5, 1).mean(dim=-1).numpy().astype(np.float32))
        g2 = torch.Generator().manual_seed(4321 + seed0 + u)
        noise.append(torch.randn(frames * steps_per_frame, generator=g2).numpy().astype(np.float32))
    return mels, noise


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe): NVML every 10 ms when the
    binding is importable, else `nvidia-smi --query-gpu` every 200 ms."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []                          # [sm_mhz, sm_max_mhz, {reasons}]
        self.stop_flag = threading.Event()
        self.source = "nvidia-smi"

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.source = "nvml"
        while not self.stop_flag.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.rows.append([sm, mx, {k for k, b in bits.items() if r & b}])
            self.stop_flag.wait(0.01)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [x.strip() for x in out.split(",")]
                if len(c) >= 7 and c[0].replace(".", "").isdigit():
                    self.rows.append([float(c[0]), float(c[1]),
                                      {self.NAMES[i] for i in range(4) if c[3 + i].lower().startswith("active")}])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = [r[0] for r in self.rows]
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm          # idle-clock samples before / after the loop do not count
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": sorted(set().union(*[r[2] for r in self.rows])), "samples": len(self.rows), "source": self.source}


def wavenet_flops_per_step(plan):
    """Algorithmic FLOPs of one WaveNet time step, un-padded channels, dilated convs + res/skip 1x1."""
    wn = plan.wavenet
    c, k, L = wn.c, wn.k, wn.n_layers
    return (L - 1) * (2 * k * c * 2 * c + 2 * c * 2 * c) + (2 * k * c * 2 * c + 2 * c * c)


def cpu_oracle_throughput(model_id, frames, sample_batch, repeats, threads):
    """audio-s/s of the restated reference forward (torch-CPU fp32) on `sample_batch` utterances."""
    import torch
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import OracleMBExWN
    torch.set_num_threads(threads)
    hp = read_config(get_config_file(model_id))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))
    orc = OracleMBExWN(hp, w, torch.float32)
    mels, noise = synthetic_batch(sample_batch, frames, plan.steps_per_frame)
    mel = np.stack(mels)
    nz = np.stack(noise)[:, :, None]
    orc.forward(mel[:1], nz[:1], return_taps=False)             # warm-up
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.forward(mel, nz, return_taps=False)
        times.append(time.perf_counter() - t0)
    audio_s = sample_batch * frames * plan.hop / plan.sample_rate
    return audio_s / float(np.median(times)), float(np.median(times))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (restated oracle; TF is not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model_id, batch, frames, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = 4 if frames <= 400 else 2
    # each step = one bounded sample of the workload
    import torch
    from mbexwn_vocoder_b200 import get_config_file, weights as W
    from mbexwn_vocoder_b200.config import read_config
    from mbexwn_vocoder_b200.plan import build_plan
    from oracle.forward import OracleMBExWN
    torch.set_num_threads(threads)
    hp = read_config(get_config_file(model_id))
    plan = build_plan(hp)
    w = W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))
    orc = OracleMBExWN(hp, w, torch.float32)
    mels, noise = synthetic_batch(sample, frames, plan.steps_per_frame)
    mel, nz = np.stack(mels), np.stack(noise)[:, :, None]
    for _ in range(max(1, min(args.warmup, 2))):
        orc.forward(mel[:1], nz[:1], return_taps=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.forward(mel, nz, return_taps=False)
    dt = time.perf_counter() - t0
    audio_s = sample * frames * plan.hop / plan.sample_rate
    value = audio_s * args.steps / dt
    sample_desc = (f"{sample} of {batch} utterances x {frames} frames per step (linear in the utterances), restated reference "
                   f"forward (torch-CPU fp32, {threads} threads), not TensorFlow")
    line = {
        "impl": "reference", "metric": "audio-sec generated/sec (24 kHz)", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, max(1, args.gpus), batch, frames, desc),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_config4(args, rank, world, local, dist):
    """BASELINE.json configs[3]: MW-SI-FD, 8192 synthetic utterances of 1-30 s (T ~ U{80..2400} frames, seed 1), LPT-sharded
    over the ranks.  Strong scaling: the set is fixed, rank r computes shard r (multi_gpu.run_shard: batches under a frame
    budget, two pinned buffer sets, host scatter / gather overlapped with the kernels); the timing covers host preparation,
    copies, kernels and the gather of every waveform into one host buffer per rank plus the ranks' result tables on rank 0."""
    import torch
    from mbexwn_vocoder_b200 import sched
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    from mbexwn_vocoder_b200.multi_gpu import imbalance, run_shard, shard_capacity
    inv = MELInverter(CONFIG4["model"], device=local, precision=args.precision, allow_synthetic_weights=True)
    eng, plan = inv.model, inv.plan
    eng.set_option("debug_taps", 0)
    eng.set_option("tc_cta_group", args.cta_group)
    eng.set_option("tc_fused", args.fused)
    rng = np.random.default_rng(CONFIG4["seed"])
    lengths = rng.integers(CONFIG4["min_frames"], CONFIG4["max_frames"] + 1, size=args.config4_utts)
    shards = sched.lpt_shards(lengths, world)
    mine = shards[rank]
    # utterance u = a window of a periodically extended base mel (35 h of distinct mels would be 3 GB of host memory)
    base = synthetic_batch(1, CONFIG4["max_frames"], plan.steps_per_frame)[0][0]
    base2 = np.concatenate([base, base])
    get_mel = lambda u: base2[(-int(u)) % CONFIG4["max_frames"]:(-int(u)) % CONFIG4["max_frames"] + int(lengths[u])]
    # the gathered result: one pre-faulted host buffer per rank, a waveform = a slice of it
    offs = np.concatenate(([0], np.cumsum([int(lengths[u]) * plan.hop for u in mine]))).astype(np.int64)
    result = np.empty(int(offs[-1]), dtype=np.float32)
    result.fill(0.0)                            # np.zeros maps untouched zero pages: fault them in here, not inside the timing
    where = {int(u): k for k, u in enumerate(mine)}

    def sink(u, w):                             # the host gather: the pinned grid's slice lands in the result buffer
        k = where[int(u)]
        return result[offs[k]:offs[k + 1]]
    # warm-up: a few batches of this rank's shard (allocates the two buffer sets and the workspace)
    warm = mine[:min(len(mine), 8)]
    host_threads = max(1, min(4, (os.cpu_count() or 4) // max(1, world)))
    cap = shard_capacity(lengths, mine, eng.halo, args.max_batch_frames)     # the warm-up allocates the buffer sets of the real run
    run_shard(inv, get_mel, lengths, warm, {}, args.max_batch_frames, seed=7, keep=False, host_threads=host_threads, capacity=cap)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = run_shard(inv, get_mel, lengths, mine, sink, args.max_batch_frames, seed=7, host_threads=host_threads, capacity=cap)
    table = np.array([[u, offs[k], offs[k + 1] - offs[k], float(np.abs(result[offs[k]:offs[k + 1]:997]).sum())]
                      for k, u in enumerate(mine)], dtype=np.float64)
    local_s = time.perf_counter() - t0
    tables = [table]
    if dist is not None:
        tables = [None] * world
        dist.all_gather_object(tables, table)   # the ranks' result tables (id, offset, samples, checksum): the final host gather
        torch.cuda.synchronize()
    wall_s = time.perf_counter() - t0
    stats = [st.wall_s, st.host_prep_s, st.host_gather_s, st.wait_s, local_s, wall_s, float(st.frames), float(st.range_reruns)]
    if dist is not None:
        t = torch.tensor(stats, dtype=torch.float64, device="cuda")
        allst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allst, t)
        allst = [x.cpu().numpy() for x in allst]
    else:
        allst = [np.array(stats)]
    wall = max(float(x[5]) for x in allst)
    n_done = sum(len(tb) for tb in tables)
    audio_s = float(sum(lengths)) * plan.hop / plan.sample_rate
    inv.model.close()
    del inv
    torch.cuda.empty_cache()
    return {"workload": f"config4: {CONFIG4['desc']}, LPT over {world} rank(s), batches <= {args.max_batch_frames} frames",
            "scaling": "strong", "utterances": int(n_done), "audio_s": audio_s, "wall_s": wall, "value": audio_s / wall,
            "unit": "audio-s/s", "precision": args.precision,
            "lpt_imbalance": imbalance(lengths, shards), "host_threads_per_rank": host_threads,
            "shard_audio_s": [float(sum(int(lengths[u]) for u in sh)) * plan.hop / plan.sample_rate for sh in shards],
            "per_rank": [{"wall_s": float(x[0]), "host_scatter_s": float(x[1]), "host_gather_s": float(x[2]),
                          "host_waiting_for_gpu_s": float(x[3]), "range_reruns": int(x[7])} for x in allst],
            "gathered": "every waveform in one host buffer per rank (node-local), result tables (id, offset, samples, checksum) "
                        "of all ranks on every rank",
            "checksum": float(sum(tb[:, 3].sum() for tb in tables))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None, choices=["f16f8", "bf16x3", "bf16", "fp32"],
                    help="default: f16f8 (fp32-accurate) for config1/2, bf16 for config3 (BASELINE.json configs)")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 sub-record (8192 mixed-length utterances, LPT-sharded)")
    ap.add_argument("--config4-utts", type=int, default=CONFIG4["utts"])
    ap.add_argument("--max-batch-frames", type=int, default=32768)
    ap.add_argument("--tc-debug", type=int, default=0, help="kernel timing experiments (invalid results): see GemmParams::debug")
    ap.add_argument("--cta-group", type=int, default=2, choices=[1, 2], help="tcgen05 tiles per CTA (1) or per CTA pair (2)")
    ap.add_argument("--fused", type=int, default=1, choices=[0, 1, 2],
                    help="WaveNet layer as one persistent kernel: 0 never (gate + res/skip launches), 1 whenever supported (default), 2 large batches only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.precision is None:
        args.precision = "bf16" if args.workload == "config3" else "f16f8"
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mbexwn_vocoder_b200.mel_inverter import MELInverter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    model_id, batch, frames, desc = WORKLOADS[args.workload]
    inv = MELInverter(model_id, device=local, precision=args.precision, allow_synthetic_weights=True)
    eng, plan = inv.model, inv.plan
    eng.set_option("debug_taps", 0)
    eng.set_option("stage_timing", 0)
    eng.set_option("tc_cta_group", args.cta_group)
    eng.set_option("tc_fused", args.fused)
    if args.tc_debug:
        eng.set_option("tc_debug", args.tc_debug)
    mels, noise = synthetic_batch(batch, frames, plan.steps_per_frame, seed0=rank * batch)
    pb = eng.prepare([frames] * batch, precision=args.precision, with_noise=True)
    pb.load(mels, noise)
    pb.upload()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------
    for _ in range(args.warmup):
        pb.run_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    ev0.record()
    for _ in range(args.steps):
        pb.run_device()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = pb.launches() * args.steps
    ws_bytes = pb.ws_bytes
    # per-stage / per-launch device times over a few more steps (outside the headline timing): CUDA events recorded
    # inside the library on the launch stream at the stage boundaries and around every WaveNet tap-GEMM launch
    eng.set_option("stage_timing", 1)
    n_avg = 5
    wn_launch = {"gate": 0.0, "resskip": 0.0, "layers": plan.wavenet.n_layers}
    for _ in range(n_avg):
        pb.run_device()
        for k, v in pb.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / n_avg
        if args.precision != "fp32":
            t = pb.wavenet_launch_ms()
            wn_launch["gate"] += t["gate"] / n_avg
            wn_launch["resskip"] += t["resskip"] / n_avg
    eng.set_option("stage_timing", 0)
    torch.cuda.synchronize()

    # ---- end to end ----------------------------------------------------------------------------------------
    # (a) C-ABI, serial: H2D, forward, D2H and a stream synchronisation inside every mbexwn_forward_host call
    for _ in range(2):
        pb.run_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pb.run_host()
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t0
    barrier()
    # (b) C-ABI, pipelined (mbexwn_forward_host_begin/_wait): two pre-filled pinned buffer sets alternate, the copies of step
    #     i +- 1 run under the kernels of step i; every step still moves its own inputs and its own waveform
    pb2 = eng.prepare([frames] * batch, precision=args.precision, with_noise=True)
    pb2.load(mels, noise)
    slots = (pb, pb2)
    for i in range(4):
        slots[i & 1].wait_host(i & 1)
        slots[i & 1].begin_host(i & 1)
    slots[0].wait_host(0)
    slots[1].wait_host(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        slots[i & 1].wait_host(i & 1)
        slots[i & 1].begin_host(i & 1)
    slots[0].wait_host(0)
    slots[1].wait_host(1)
    torch.cuda.synchronize()
    e2e_cabi_s = time.perf_counter() - t0
    barrier()
    del pb2
    # (c) the product API, NumPy in -> NumPy out: MELInverter.synth_stream over the steps' batches -- layout scatter into
    #     pinned memory, H2D, forward (noise channel drawn in-kernel like the reference draws it inside the graph), D2H and the
    #     copies of the waveforms out of the pinned grid are all inside the timing
    def api_steps(n):
        got = 0
        for waves in inv.synth_stream(mels for _ in range(n)):
            got += sum(w.size for w in waves)
        return got
    api_steps(3)
    barrier()
    t0 = time.perf_counter()
    n_samples = api_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    assert n_samples == args.steps * batch * frames * plan.hop
    api_h2d = (batch * frames + eng.halo * (batch + 1)) * plan.mel_channels * 4
    api_d2h = (batch * frames + eng.halo * (batch + 1)) * plan.hop * 4
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e_serial_s, e2e_cabi_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_serial_s, e2e_cabi_s = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    # ---- BASELINE.json configs[3]: the mixed-length set, LPT-sharded over the ranks (strong scaling) ------------------
    config4 = None
    if not args.no_config4 and args.workload == "config2":
        config4 = run_config4(args, rank, world, local, dist if world > 1 else None)

    audio_s_step = batch * frames * plan.hop / plan.sample_rate          # per rank
    value = world * audio_s_step * args.steps / (dev_ms / 1e3)
    e2e = world * audio_s_step * args.steps / e2e_s

    pk = peaks()
    rows = batch * frames * plan.steps_per_frame
    wn = plan.wavenet
    L = wn.n_layers
    wn_flops = wavenet_flops_per_step(plan) * rows
    wn_ms = stage_acc["wavenet"]
    # executed tensor work in bf16-rate product equivalents: bf16x3 = 3 products; f16f8 = 1 fp16 product + 2 e4m3 products
    # that run at twice the rate
    factor = {"bf16x3": 3, "f16f8": 2}.get(args.precision, 1)
    peak = pk["bf16_tflops_sustained"]
    # dominant kernel: the gate tap-GEMM (dilated conv, K = k C, N = 2 C), one launch per layer
    gate_flops = 2.0 * wn.k * wn.c * 2 * wn.c * rows                     # algorithmic, un-padded, per launch
    fused_layer = args.precision != "fp32" and wn_launch["gate"] > 0 and wn_launch["resskip"] == 0.0
    if fused_layer:
        # one persistent kernel per layer: dilated conv + gate + res/skip 1x1 + residual update (k_wavenet_layer.cu)
        gate_flops = wn_flops / L                                        # algorithmic FLOPs of an average layer launch
        gate_ms = wn_launch["gate"] / L
        achieved = gate_flops / (gate_ms / 1e3) / 1e12
        kernel = "wn_layer_kernel (tcgen05: dilated-conv tap-GEMM + gate + res/skip 1x1 + residual update, one launch per layer)"
    elif args.precision != "fp32" and wn_launch["gate"] > 0:
        gate_ms = wn_launch["gate"] / L
        achieved = gate_flops / (gate_ms / 1e3) / 1e12
        kernel = "wn_gemm_kernel<EPI_GATE> (tcgen05 tap-GEMM of the dilated conv + tanh*sigmoid gate epilogue)"
    else:
        gate_ms = wn_ms / L
        achieved = wn_flops / L / (gate_ms / 1e3) / 1e12
        kernel = "fp32 SIMT WaveNet layer (conv1d + gate + res/skip kernels)"
    traffic = ncu_traffic(args.workload, args.precision, "wn_layer_kernel" if fused_layer else "wn_gemm_kernel<EPI_GATE>")
    roofline = {
        "bound": "tensor", "kernel": kernel,
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": f"{pk['source']} bf16 dense sustained (MEASURED_PEAKS.json; kernel timed inside a long step)",
        "executed_flop_factor": factor, "frac_executed": achieved * factor / peak,
        "launches_per_step": L, "avg_launch_ms": gate_ms, "algorithmic_flops_per_launch": gate_flops,
        "traffic": traffic,
        "wavenet_stage": {"ms": wn_ms, "gate_ms": wn_launch["gate"], "resskip_ms": wn_launch["resskip"],
                          "algorithmic_tflops": wn_flops / (wn_ms / 1e3) / 1e12,
                          "frac_executed": wn_flops * factor / (wn_ms / 1e3) / 1e12 / peak},
    }
    # HBM-bound stages: algorithmic bytes per audio-second
    audio_s = audio_s_step
    exc_bytes = audio_s * (plan.sample_rate / plan.pulse_rate_factor * 4 + 1600 * 4 + 1600 * plan.wavenet.c_in * 4)
    syn_bytes = audio_s * (1600 * plan.subbands * 4 + 80 * plan.n_ceps * 4 + plan.sample_rate * 4)
    stages = {k: {"ms": v} for k, v in stage_acc.items()}
    stages["excitation"].update({"algorithmic_gbs": exc_bytes / (stage_acc["excitation"] / 1e3) / 1e9,
                                 "frac_hbm": exc_bytes / (stage_acc["excitation"] / 1e3) / 1e9 / pk["hbm_gbs"]})
    syn_ms = stage_acc["post_pqmf"] + stage_acc["stft_ola"]
    stages["synthesis(post_pqmf+stft_ola)"] = {"ms": syn_ms, "algorithmic_gbs": syn_bytes / (syn_ms / 1e3) / 1e9,
                                               "frac_hbm": syn_bytes / (syn_ms / 1e3) / 1e9 / pk["hbm_gbs"]}

    line = {
        "metric": "audio-sec generated/sec (24 kHz)", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"f16f8": "f16+2xe4m3 split (fp16 product + two e4m3 correction products, fp32 accumulate; fp32-accurate)",
                  "bf16x3": "bf16x3 (bf16 hi/lo split, fp32 accumulate; fp32-accurate)", "bf16": "bf16",
                  "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": config_dict(args, world, batch, frames, desc),
        "kernel_options": {"tc_cta_group": args.cta_group, "fused_layer_kernel": bool(fused_layer),
                           "l2": (f"per-step working set {ws_bytes / 1e6:.0f} MB of activations "
                                  + (">> 126 MB L2; no flush needed" if ws_bytes > 4 * 126e6 else "(comparable to the 126 MB L2: not "
                                     "flushed, treat as an L2-warm figure)"))},
        "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": api_h2d, "d2h_bytes_per_step": api_d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "call": "MELInverter.synth_stream: list of NumPy mels in, list of NumPy waveforms out (scatter into pinned memory, "
                        "H2D, forward, D2H and the copies out inside the timing; two buffer sets)",
                "cabi": {"value": world * audio_s_step * args.steps / e2e_cabi_s, "ms_per_step": 1e3 * e2e_cabi_s / args.steps,
                         "call": "mbexwn_forward_host_begin/_wait on pre-filled pinned buffers (explicit noise input)"},
                "serial": {"value": world * audio_s_step * args.steps / e2e_serial_s, "ms_per_step": 1e3 * e2e_serial_s / args.steps,
                           "call": "mbexwn_forward_host (H2D, forward, D2H, synchronise per call)"}},
        "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "stages": stages,
    }
    if config4 is not None:
        line["config4"] = config4

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = 4 if frames <= 400 else 2
        v, med = cpu_oracle_throughput(model_id, frames, sample, 2, threads)
        # also at 1 thread (the README's "single laptop core" claim) and 2 threads (the reference CLI default,
        # bin/resynth_mel.py:120), one utterance each
        sweep = {str(threads): v}
        for n in (1, 2):
            if n < threads:
                sweep[str(n)] = cpu_oracle_throughput(model_id, frames, 1, 1, n)[0]
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                "sample": f"{sample} of {batch} utterances x {frames} frames, median of 2 runs "
                                          f"({med:.1f} s each), restated reference forward (torch-CPU fp32), not TensorFlow",
                                "by_threads": sweep}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
