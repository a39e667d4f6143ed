#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched ship-in-transit environment on N B200s.

One bench "step" = one full episode of every environment of the workload: reset() followed by
max_sampling_frequency (9) step(action) calls, each of which runs its data-dependent number of
simulator steps (_step(): both ships of the pair + termination/reward evaluation).  The metric
counts simulator steps actually integrated (device counters), not idle lanes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--workload colav_iw|rl]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...       # the CPU implementation on the box's host cores

Workload at N = 1 (BASELINE.json configs[1]): the run_simplified_IW_model.py test+obs SimpleShipModel
pair with HeadingBySampledRouteController, batched to 1e5 environments per GPU, dt = 4 s, per-env
scoping angles ~ U(-pi/6, pi/6) (torch.Generator seed 0) and +-50 m start-position jitter (seed 1;
50 m keeps every ship inside the map horizon at t = 0 -- the obstacle ship starts 100 m from the edge).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/s"
N_RL_STEPS = 9

# FP64 work per simulator step of one environment (both ships + evaluation), see DESIGN.md section 5:
#  - algorithmic count from SURVEY.md section 8(d) (formula-level add/mul/div/sqrt, FMA = 2, the 26/31
#    transcendental calls counted as 1 each);
#  - executed count = FP64 flops issued per env-step by the kernel (DFMA x2 + DADD + DMUL), from the
#    ncu capture under profiles/ (includes the CUDA math library's sin/cos/atan2/exp internals and
#    the map-geometry tests); filled in from profiles/r01_ncu_summary.md.
FLOP_ALGO = {"colav_iw": 370.0 + 26.0, "rl": 450.0 + 31.0}
FLOP_EXEC = {"colav_iw": 561.0, "rl": 976.0}     # profiles/r01_ncu_summary.md part 4, section 2 (fast build)
# DRAM bytes (read + written) of one k_env<MODE_STEP> launch over 1e5 environments, from the ncu --set full
# capture summarised in profiles/r01_ncu_summary.md part 2 (dram__bytes_read.sum + dram__bytes_write.sum)
TRAFFIC_PER_LAUNCH_1E5 = {"colav_iw": 40.3e6, "rl": 40.7e6}
# HBM bytes per env-step when every simulator step is its own launch (K = 1): DESIGN.md section 4
BYTES_K1 = 2 * 2 * (17 * 8 + 4) + 2 * (5 * 8 + 2 * 4) + 32 + 8 + 4 + 4     # ship rows r+w, env rows r+w, outputs = 704 B (ABI v5)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=100_000, help="environments per GPU")
    ap.add_argument("--workload", default="colav_iw", choices=["colav_iw", "rl"])
    ap.add_argument("--collav", default="none", choices=["none", "simple", "sbmpc"],
                    help="collision avoidance of the ship under test (the headline workload uses 'none')")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"], help="device code build (see DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--k1-launches", type=int, default=64, help="launches of the one-step-per-launch probe")
    return ap.parse_args()


def make_inputs(workload, envs, rank, collav="none"):
    """Synthetic inputs of the workload: actions [B, 9] float64 and jittered initial states."""
    import torch
    from ast_sac_b200 import scenarios as S
    args = S.get_env_args(time_step=4, collav_mode=collav)
    if workload == "rl":
        assets, m = S.build_rl_assets(args)
    else:
        assets, m = S.build_colav_assets(args, iw=True)
    gen = torch.Generator().manual_seed(0 + 7919 * rank)
    actions = (torch.rand((envs, N_RL_STEPS), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    init = S.jittered_init_states(assets, envs, pos_jitter_m=50.0, seed=1 + 7919 * rank, device="cpu")
    return args, assets, m, actions, init


def workload_name(workload, envs, collav="none"):
    suffix = "" if collav == "none" else f", collav_mode={collav}"
    if workload == "rl":
        return (f"rl_env MultiShipRLEnv: ShipModelAST (PTI) test+obs pair, full AST reward + map, {envs} envs/GPU, "
                f"dt=4, 9 step() per episode{suffix}")
    return (f"run_colav MultiShipEnv (run_simplified_IW_model.py): SimpleShipModel test+obs pair + "
            f"HeadingBySampledRouteController, {envs} envs/GPU, dt=4, 9 step() per episode{suffix}")


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, via NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _init(self):
        """NVML handle and the reason-bit names; done before the timed region starts (nvmlInit takes ~100 ms).
        Sampled every 15 ms (every rank samples its own GPU; NVML queries are not free, so the period stays coarse)."""
        import pynvml as nv
        nv.nvmlInit()
        self._nv = nv
        self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        self._names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }

    def _sample(self):
        nv, h = self._nv, self._h
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, name in self._names.items():
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        try:
            while not self._stop.is_set():
                self._sample()
                time.sleep(0.015)
        except Exception as exc:
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def start(self):
        try:
            self._init()
        except Exception as exc:   # NVML missing: record that instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        if self._t:
            try:
                self._sample()                   # at least one sample while the GPU is still under load
            except Exception:
                pass
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_episode_rate(workload, n_episodes, threads=0, seed_rank=0, collav="none"):
    """Time the CPU oracle on a bounded sample of the same workload; returns (env-steps/s, steps, secs)."""
    from oracle import oracle as O
    args, assets, m, actions, init = make_inputs(workload, n_episodes, seed_rank, collav)
    kind = O.ENV_RL if workload == "rl" else O.ENV_COLAV_IW
    cfg = O.env_config_from_assets(assets, m, args, kind)
    init_np = init.numpy().reshape(7, n_episodes, 2)
    base = np.array([[assets[r].ship_model.simulation_config.initial_north_position_m,
                      assets[r].ship_model.simulation_config.initial_east_position_m] for r in range(2)])
    jitter = np.stack([init_np[0] - base[:, 0][None, :], init_np[1] - base[:, 1][None, :]], axis=-1)  # [B, 2, 2]
    t0 = time.perf_counter()
    total, _, _ = O.bench_episodes(cfg, actions.numpy(), jitter_ne=jitter, n_threads=threads)
    dt = time.perf_counter() - t0
    return total / dt, total, dt


def calibrated_cpu_sample(workload, target_s=12.0, collav="none"):
    cores = os.cpu_count() or 1
    rate, steps, secs = cpu_episode_rate(workload, 64 * cores, collav=collav)
    per_episode = secs / (64 * cores)
    n = int(max(64 * cores, min(400_000, target_s / max(per_episode, 1e-9))))
    return n, cores


def run_reference(a, rank, world):
    if rank != 0:
        return
    n, cores = calibrated_cpu_sample(a.workload, target_s=10.0, collav=a.collav)
    for _ in range(a.warmup):
        cpu_episode_rate(a.workload, max(64, n // 8), collav=a.collav)
    tot_steps, tot_s = 0, 0.0
    for _ in range(a.steps):
        _, steps, secs = cpu_episode_rate(a.workload, n, collav=a.collav)
        tot_steps += steps
        tot_s += secs
    value = tot_steps / tot_s
    sample = f"{n} episodes per step (reset + up to 9 step() each) of the same workload, all host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot_s / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.workload, a.envs, a.collav), "sample": sample},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference simulator (oracle/), pthreads over episodes; the "
                                 "reference itself is single-process Python (~1.4e3 env-steps/s on one core, "
                                 "BASELINE.md section 2) and cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from ast_sac_b200 import _lib as L
    from ast_sac_b200 import parallel as PAR
    from ast_sac_b200 import scenarios as S

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    B = a.envs
    args, assets, m, actions_cpu, init_cpu = make_inputs(a.workload, B, rank, a.collav)
    init_dev = init_cpu.to(dev)
    actions_dev = actions_cpu.to(dev).t().contiguous()                  # [9, B]: one contiguous row per step() call
    actions_host = np.ascontiguousarray(actions_cpu.numpy().T)          # [9, B] rows for the host API
    if a.workload == "rl":
        env = S.MultiShipRLEnv(assets=assets, map=m, args=args, num_envs=B, device=dev, init_states=init_dev,
                               math_mode=a.math)
    else:
        env = S.MultiShipEnv(assets=assets, map=m, args=args, num_envs=B, device=dev, init_states=init_dev,
                             math_mode=a.math)
    fp64_peak = L.measure_fp64_peak(local_rank, repeats=5)

    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    ep_return = torch.zeros(B, dtype=torch.float64, device=dev)
    ep_rl = torch.zeros(B, dtype=torch.int32, device=dev)
    ep_steps = torch.zeros(B, dtype=torch.int32, device=dev)

    launch_steps = torch.zeros(1 + N_RL_STEPS, dtype=torch.int64, device=dev)

    def one_episode(events=None):
        """reset + 9 step() on device-resident actions; returns list of (start, end) event pairs."""
        ep_return.zero_()
        ep_rl.zero_()
        ep_steps.zero_()
        launch_steps.zero_()
        pairs = []
        for j in range(-1, N_RL_STEPS):
            if events is not None:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
            if j < 0:
                env.reset()
            else:
                env.step(actions_dev[j])
            if events is not None:
                e.record(stream)
                pairs.append((s, e))
            if j >= 0:
                ep_return.add_(env.reward_buf)
                ep_rl.add_((env.nsub_buf > 0).to(torch.int32))
                ep_steps.add_(env.nsub_buf)
                launch_steps[j + 1] = env.nsub_buf.sum()
        return pairs

    for _ in range(a.warmup):
        one_episode()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    c0 = env.total_substeps()
    step_ms, launch_ms, kernel_ms = [], [], []
    lib = L.load()
    L.check(lib.shipenv_time_env_kernel(env._handle, 1))   # CUDA events around k_env itself, on its own stream
    barrier()
    t_wall0 = time.perf_counter()
    import ctypes
    for _ in range(a.steps):
        l2_flush.fill_(1.0)                      # > 126 MB L2: evict the state between timed iterations
        pairs = one_episode(events=True)
        torch.cuda.synchronize(dev)
        ms = [s.elapsed_time(e) for s, e in pairs]
        launch_ms.append(ms)
        step_ms.append(sum(ms))
        kms = ctypes.c_double()
        L.check(lib.shipenv_env_kernel_ms(env._handle, ctypes.byref(kms)))
        kernel_ms.append(kms.value)
    barrier()
    L.check(lib.shipenv_time_env_kernel(env._handle, 0))
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    steps_done = env.total_substeps() - c0
    local_time_s = sum(step_ms) / 1e3

    # ---- per-launch roofline of the dominant kernel (k_env<.., MODE_STEP>): the 9 step() launches
    launch_ms = np.array(launch_ms)                       # [K, 10]: reset() + 9 step() calls (prologue + k_env each)
    step_kernel_ms = np.array(kernel_ms)                  # the 9 k_env launches of each episode, device time
    env_steps_per_episode = steps_done / a.steps
    flops_exec = FLOP_EXEC[a.workload] * env_steps_per_episode
    achieved_tf = flops_exec / (step_kernel_ms.mean() * 1e-3) / 1e12
    if a.collav == "sbmpc":
        # the flop counts above are those of the plain kernels; an SBMPC evaluation adds 1e3 ... 3e4 flop to the
        # steps it is active in (profiles/r01_ncu_summary.md), so no per-step constant applies
        achieved_tf = float("nan")
    roofline = {
        "bound": "fp64", "kernel": "k_env<MODE_STEP> (9 launches per episode)",
        "ms_per_launch": float(step_kernel_ms.mean() / N_RL_STEPS),
        "achieved": None if achieved_tf != achieved_tf else achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": None if achieved_tf != achieved_tf else achieved_tf / fp64_peak,
        "peak_source": "DFMA microbenchmark measured live in this run (MEASURED_PEAKS.json has no FP64 entry)",
        "flop_per_env_step_executed": FLOP_EXEC[a.workload], "flop_per_env_step_algorithmic": FLOP_ALGO[a.workload],
        "achieved_algorithmic": FLOP_ALGO[a.workload] * env_steps_per_episode / (step_kernel_ms.mean() * 1e-3) / 1e12,
        "traffic": TRAFFIC_PER_LAUNCH_1E5[a.workload] if (B == 100_000 and a.collav == "none") else None,
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/)",
        "algorithmic_bytes_per_launch": float(B * (2 * (17 * 8 + 4) * 2 + 2 * (5 * 8 + 8) + 48)),
        "share_of_step": float(step_kernel_ms.sum() / launch_ms.sum()),
    }

    # ---- HBM roofline of the one-simulator-step-per-launch configuration (K = 1, HBM-bound)
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json (burst)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    env.reset()
    torch.cuda.synchronize(dev)
    k1_ms = []
    for _ in range(a.k1_launches):
        l2_flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        # the C ABI entry point itself (env._step() adds torch ops that build the info tensors)
        L.check(L.load().shipenv_substeps(env._handle, 1, env._stream_ptr()))
        e.record(stream)
        torch.cuda.synchronize(dev)
        k1_ms.append(s.elapsed_time(e))
    k1 = float(np.mean(k1_ms[4:])) if len(k1_ms) > 8 else float(np.mean(k1_ms))
    k1_gbs = BYTES_K1 * B / (k1 * 1e-3) / 1e9
    roofline_k1 = {"bound": "hbm", "kernel": "k_env<MODE_SUBSTEPS>, k=1 (one _step() per launch)",
                   "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                   "peak_source": hbm_src, "bytes_per_env_step": BYTES_K1, "ms_per_launch": k1,
                   "env_steps_per_s": B / (k1 * 1e-3), "traffic": None}

    # ---- e2e: the reference-facing host API (numpy in / numpy out), H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        def host_episode():
            env.reset_host()
            n = 0
            for j in range(N_RL_STEPS):
                _, _, _, nsub = env.step_host(actions_host[j])
                n += int(nsub.sum())
            return n
        host_episode()
        barrier()
        t0 = time.perf_counter()
        n_e2e = 0
        for _ in range(a.steps):
            n_e2e += host_episode()
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e_local = torch.tensor([n_e2e, t_e2e], dtype=torch.float64, device=dev)
    else:
        e2e_local = torch.zeros(2, dtype=torch.float64, device=dev)

    # ---- aggregate over ranks: sum of steps, max of time; NCCL all-gather of the episode statistics
    agg = torch.tensor([steps_done, local_time_s, t_wall], dtype=torch.float64, device=dev)
    one_episode()        # a last untimed episode leaves complete per-env statistics in the buffers
    stats = PAR.episode_stats(env.info_buf & L.INFO_EVENT_MASK, ep_return, ep_steps, ep_rl)
    if world > 1:
        allagg = torch.empty(world * 3, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allagg, agg)
        allagg = allagg.view(world, 3)
        alle2e = torch.empty(world * 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(alle2e, e2e_local)
        alle2e = alle2e.view(world, 2)
    else:
        allagg, alle2e = agg.view(1, 3), e2e_local.view(1, 2)
    gathered = PAR.gather_stats(stats)
    total_steps = float(allagg[:, 0].sum())
    max_time = float(allagg[:, 1].max())
    value = total_steps / max_time

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * max_time / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.workload, B, a.collav), "envs_per_gpu": B, "envs_total": B * world,
                       "math_mode": a.math,
                       "sharding": "contiguous env blocks per rank, no per-step collective",
                       "l2": "256 MB buffer written between timed iterations (state < 126 MB L2)",
                       "env_steps_per_episode_mean": env_steps_per_episode / B,
                       "wall_s_timed_region": float(allagg[:, 2].max())},
            "clocks": clocks,
            "gpu_launches": int(a.steps * (1 + 2 * N_RL_STEPS)),     # k_reset + 9 x (k_prologue + k_env) per episode
            "launch_ms_mean": [round(float(x), 4) for x in launch_ms.mean(axis=0)],
            "launch_env_steps": [int(x) for x in launch_steps.tolist()],
            "roofline": roofline, "roofline_hbm_k1": roofline_k1,
            "episode_stats": PAR.summarise(gathered),
        }
        if not a.no_e2e:
            e2e_steps, e2e_time = float(alle2e[:, 0].sum()), float(alle2e[:, 1].max())
            line["e2e"] = {"value": e2e_steps / e2e_time, "unit": METRIC,
                           "h2d_bytes_per_step": int(N_RL_STEPS * B * 8),
                           "d2h_bytes_per_step": int(N_RL_STEPS * B * (32 + 8 + 4 + 4) + B * 32),
                           "api": "reset_host() + 9 x step_host(): numpy actions in, numpy obs/reward/info out "
                                  "through shipenv_reset_host / shipenv_step_host (pinned staging inside the C ABI)"}
        if world == 1 and not a.no_cpu_baseline:
            n, cores = calibrated_cpu_sample(a.workload, target_s=12.0, collav=a.collav)
            rate, steps, secs = cpu_episode_rate(a.workload, n, collav=a.collav)
            line["cpu_baseline"] = {"value": rate, "unit": METRIC, "cores": cores, "kind": "port",
                                    "sample": f"{n} episodes of the same workload ({steps} env-steps, {secs:.1f} s), "
                                              f"oracle/shipsim_oracle.c on {cores} host threads"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    run_b200(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
