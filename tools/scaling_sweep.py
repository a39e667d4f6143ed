#!/usr/bin/env python
"""BASELINE config 4: env-steps/s for 1e3 ... 1e7 environments, env-sharded over the ranks of the launch
(one process per GPU, `env b -> rank floor(b * G / B)`, no per-step collective), in two regimes:

  episode   reset() + 9 x step(action): data-dependent episode lengths, state in registers for a whole call
  substeps  k x _step() per launch for k in {1, 16, 128}: k = 1 is the HBM-bound configuration (648 B of
            state per env-step cross HBM), large k the FP64-bound one

    python tools/scaling_sweep.py [--workload colav_iw|rl] [--envs 1000,10000,...]
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/scaling_sweep.py ...

One JSON line per point (rank 0): total envs, ranks, regime, env-steps/s (sum of steps / max-over-ranks device
time), and the FP64 / HBM roofline fractions computed like bench.py.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402
from ast_sac_b200 import _lib as L  # noqa: E402
from ast_sac_b200 import parallel as PAR  # noqa: E402
from ast_sac_b200 import scenarios as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="colav_iw", choices=["colav_iw", "rl"])
    ap.add_argument("--envs", default="1000,10000,100000,1000000,10000000")
    ap.add_argument("--ks", default="1,16,128")
    ap.add_argument("--substeps-total", type=int, default=256, help="_step() calls per environment in the substeps regime")
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--regimes", default="episode,substeps", help="comma list of: episode, substeps")
    ap.add_argument("--no-warmup", action="store_true", help="single measured pass per point (ncu captures)")
    a = ap.parse_args()
    regimes = a.regimes.split(",")
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = L.measure_fp64_peak(local, repeats=3)
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def reduce(steps, ms):
        t = torch.tensor([steps, ms], dtype=torch.float64, device=dev)
        if world > 1:
            out = torch.empty(world * 2, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, t)
            out = out.view(world, 2)
            return float(out[:, 0].sum()), float(out[:, 1].max())
        return steps, ms

    for total in [int(x) for x in a.envs.split(",")]:
        lo, hi = PAR.shard_range(total, rank, world)
        B = hi - lo
        if B <= 0:
            raise SystemExit("fewer environments than ranks")
        args, assets, m, actions, init = BN.make_inputs(a.workload, B, rank)
        actions, init = actions.to(dev), init.to(dev)
        cls = S.MultiShipRLEnv if a.workload == "rl" else S.MultiShipEnv
        env = cls(assets=assets, map=m, args=args, num_envs=B, device=dev, init_states=init)
        ev = lambda: torch.cuda.Event(enable_timing=True)

        def episode():
            env.reset()
            for j in range(BN.N_RL_STEPS):
                env.step(actions[:, j])

        counts, stale = BN.kernel_counts(a.workload, "none")
        flop_exec = None if (counts is None or stale) else counts["flop_exec_per_env_step"]
        if not a.no_warmup or "episode" in regimes:
            episode()                                          # warm-up (module load, first touch)
        best = None
        for _ in range(a.repeats if "episode" in regimes else 0):
            flush.fill_(1.0)
            c0 = env.total_substeps()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            s, e = ev(), ev()
            s.record(); episode(); e.record()
            torch.cuda.synchronize()
            steps, ms = reduce(env.total_substeps() - c0, s.elapsed_time(e))
            if best is None or steps / ms > best[0] / best[1]:
                best = (steps, ms)
        if best is not None and rank == 0:
            rate = best[0] / (best[1] * 1e-3)
            tf = None if flop_exec is None else rate / world * flop_exec / 1e12
            print(json.dumps({"envs_total": total, "n_gpus": world, "regime": "episode", "workload": a.workload,
                              "env_steps_per_s": rate, "ms": best[1], "env_steps": best[0],
                              "fp64_tflops_per_gpu": tf, "fp64_frac_of_peak": None if tf is None else tf / peak,
                              "fp64_peak_tflops": peak}), flush=True)
        for k in ([int(x) for x in a.ks.split(",")] if "substeps" in regimes else []):
            n_launch = max(1, a.substeps_total // k)
            env.reset()
            if not a.no_warmup:
                env._step(k)
            best = None
            for _ in range(a.repeats):
                env.reset()
                flush.fill_(1.0)
                c0 = env.total_substeps()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                s, e = ev(), ev()
                s.record()
                for _ in range(n_launch):
                    env._step(k)
                e.record()
                torch.cuda.synchronize()
                steps, ms = reduce(env.total_substeps() - c0, s.elapsed_time(e))
                if best is None or steps / ms > best[0] / best[1]:
                    best = (steps, ms)
            rate = best[0] / (best[1] * 1e-3)
            if rank == 0:
                tf = None if flop_exec is None else rate / world * flop_exec / 1e12
                gbs = rate / world * BN.BYTES_K1 / k / 1e9
                print(json.dumps({"envs_total": total, "n_gpus": world, "regime": f"substeps k={k}", "workload": a.workload,
                                  "launches": n_launch, "env_steps_per_s": rate, "ms": best[1], "env_steps": best[0],
                                  "fp64_tflops_per_gpu": tf, "fp64_frac_of_peak": None if tf is None else tf / peak,
                                  "hbm_gbs_per_gpu": gbs, "hbm_frac_of_peak": gbs / hbm}), flush=True)
        env.close()
        del env
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
