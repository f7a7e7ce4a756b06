#!/usr/bin/env python
"""One-off runs of the BASELINE.json configurations that are not the bench line (configs[2], [3], [4]).

    python tools/run_configs.py config3            # MW-VO-FD, 256 x 10 s, bf16, 1 GPU
    python tools/run_configs.py config4 [--utts N] # MW-SI-FD, N (8192) utterances of 1-30 s, LPT-sharded over the ranks
    python tools/run_configs.py config5 [--minutes M]  # one M-minute (10) mel, chunked long-form synthesis

Under torchrun, config4 shards the utterances over the ranks (no data-path collective) and rank 0 prints the line.
Each prints one JSON line; inputs are the synthetic mels of SURVEY.md 8d.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_mel(frames, seed):
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    x = torch.clamp(torch.randn(frames + 4, 80, generator=g) * 2.0 - 4.0, min=float(np.log(1e-5)), max=2.0)
    return x.unfold(0, 5, 1).mean(dim=-1).numpy().astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["config3", "config4", "config5"])
    ap.add_argument("--utts", type=int, default=8192)
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--max-batch-frames", type=int, default=32768)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from mbexwn_vocoder_b200.mel_inverter import MELInverter
    from mbexwn_vocoder_b200 import sched

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "config3":
        prec = args.precision or "bf16"
        inv = MELInverter("VOICE", device=local, precision=prec, allow_synthetic_weights=True)
        eng, plan = inv.model, inv.plan
        eng.set_option("debug_taps", 0)
        B, T = 256, 800
        mels = [synthetic_mel(T, rank * B + u) for u in range(B)]
        pb = eng.prepare([T] * B, precision=prec, with_noise=False)
        pb.load(mels)
        pb.upload()
        for _ in range(2):
            pb.run_device()
        barrier()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            pb.run_device()
        barrier()
        dt = (time.perf_counter() - t0) / n
        audio = B * T * plan.hop / plan.sample_rate
        line = {"config": "config3: MW-VO-FD batch 256 x 10 s", "precision": prec, "n_gpus": world,
                "audio_s_per_s": world * audio / dt, "ms_per_step": 1e3 * dt, "workspace_gb": pb.ws_bytes / 1e9,
                "finite": bool(np.isfinite(pb.out_dev[:100000].cpu().numpy()).all())}
    elif args.config == "config4":
        prec = args.precision or "f16f8"
        inv = MELInverter("SING", device=local, precision=prec, allow_synthetic_weights=True)
        eng, plan = inv.model, inv.plan
        eng.set_option("debug_taps", 0)
        rng = np.random.default_rng(1)
        lengths = rng.integers(80, 2401, size=args.utts)
        shards = sched.lpt_shards(lengths, world)
        mine = shards[rank]
        # batches of utterances (in LPT order of this shard) up to max_batch_frames padded frames
        groups, cur, frames = [], [], 0
        for u in mine:
            n = int(lengths[u]) + eng.halo
            if cur and frames + n > args.max_batch_frames:
                groups.append(cur)
                cur, frames = [], 0
            cur.append(u)
            frames += n
        if cur:
            groups.append(cur)
        base = synthetic_mel(2400, 0)
        base2 = np.concatenate([base, base])                  # utterance u = a window of the periodically extended base
        off = lambda u: (-int(u)) % 2400                      # same values as np.roll(base, u)[:len], without the copy

        # two buffer sets of full capacity, re-bound to every group's geometry (no allocation inside the timed loop);
        # the host prepares group i + 1 and drains group i - 1 while the GPU works on group i
        cap_utts = max(len(g) for g in groups)
        slots = [eng.prepare([1], precision=prec, with_noise=False, capacity_frames=args.max_batch_frames + eng.halo,
                             capacity_utts=cap_utts) for _ in range(2)]
        in_flight = [None, None]
        barrier()
        t0 = time.perf_counter()
        total_frames, checksum = 0, 0.0

        def drain(s):
            nonlocal checksum
            if in_flight[s] is None:
                return
            slots[s].wait_host(s)
            n = slots[s].layout.n_frames * plan.hop
            checksum += float(np.abs(slots[s].out_host.numpy()[:n:997]).sum())
            in_flight[s] = None

        for i, grp in enumerate(groups):
            s = i & 1
            drain(s)
            pb = slots[s].rebind([int(lengths[u]) for u in grp])
            pb.set_utt_ids(grp)
            pb.load([base2[off(u):off(u) + int(lengths[u])] for u in grp])
            pb.begin_host(s, seed=7)
            in_flight[s] = grp
            total_frames += int(sum(lengths[u] for u in grp))
        drain(0)
        drain(1)
        barrier()
        dt = time.perf_counter() - t0
        tot = torch.tensor([float(total_frames), checksum, dt], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = tot.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dt = float(mx[2])
        audio = float(tot[0]) * plan.hop / plan.sample_rate
        line = {"config": f"config4: MW-SI-FD {args.utts} utterances of 1-30 s, LPT over {world} ranks", "precision": prec,
                "n_gpus": world, "audio_s": audio, "wall_s": dt, "audio_s_per_s": audio / dt, "e2e": True,
                "shard_frames": [int(sum(lengths[u] for u in s)) for s in shards], "calls_rank0": len(groups),
                "checksum": float(tot[1])}
    else:
        prec = args.precision or "f16f8"
        inv = MELInverter("SPEECH", device=local, precision=prec, allow_synthetic_weights=True)
        plan = inv.plan
        inv.model.set_option("debug_taps", 0)
        T = int(args.minutes * 60 * plan.sample_rate / plan.hop)
        mel = np.concatenate([synthetic_mel(min(2400, T - s), s) for s in range(0, T, 2400)])[:T]
        noise = np.random.default_rng(0).standard_normal(T * plan.steps_per_frame, dtype=np.float32)
        inv.synth_long_from_mel(mel[:2000], noise=noise[:2000 * plan.steps_per_frame], chunk_frames=400)     # warm-up
        torch.cuda.synchronize()
        _, cold = inv.synth_long_from_mel(mel, noise=noise, chunk_frames=400, max_batch_frames=args.max_batch_frames,
                                          return_info=True)          # first call of this length: allocates the buffers it caches
        torch.cuda.synchronize()
        out, info = inv.synth_long_from_mel(mel, noise=noise, chunk_frames=400, max_batch_frames=args.max_batch_frames,
                                            return_info=True)
        info["cold_total_s"] = cold["total_s"]
        line = {"config": f"config5: one {args.minutes:g}-minute mel, 400-frame chunks + {info['context_frames']} context frames",
                "precision": prec, "n_gpus": 1, "audio_s": info["audio_s"], "first_chunk_latency_ms": 1e3 * info["first_chunk_latency_s"],
                "f0_pass_ms": 1e3 * info["f0_pass_s"], "total_s": info["total_s"], "audio_s_per_s": info["audio_s"] / info["total_s"],
                "n_windows": info["n_windows"], "finite": bool(np.isfinite(out).all()),
                "main_pass_ms": 1e3 * info.get("main_pass_s", 0.0), "first_call_total_s": info["cold_total_s"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
