#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched ship-in-transit environment on N B200s.

One bench "step" = `--episodes-per-step` full episodes of every environment of the workload: reset() followed by
max_sampling_frequency (9) step(action) calls, each of which runs its data-dependent number of simulator steps
(_step(): both ships of the pair + termination / reward evaluation).  The metric counts simulator steps actually
integrated (device counters), not idle lanes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--workload colav_iw|rl] [--collav none|simple|sbmpc]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...       # the CPU implementation on the box's host cores

Headline workload (BASELINE.json configs[1]): the run_simplified_IW_model.py test+obs SimpleShipModel pair with
HeadingBySampledRouteController, batched to 1e5 environments per GPU, dt = 4 s, per-env scoping angles
~ U(-pi/6, pi/6) (torch.Generator seed 0) and +-50 m start-position jitter (seed 1; 50 m keeps every ship inside the
map horizon at t = 0 -- the obstacle ship starts 100 m from the edge).  At N = 1 the same JSON line also carries the
other claimed workloads under "workloads": config 3 (1e6 ShipModelAST pairs, rl env) and the two SBMPC variants at 1e5.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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
#  - executed count (DFMA x2 + DADD + DMUL per env-step) and the DRAM traffic of a launch are MEASURED values: they are
#    read from profiles/kernel_counts.json, which tools/ncu_kernel_counts.py writes from an ncu capture together with
#    the hash of the device sources it was taken on.  When the sources have changed since, the roofline block says
#    "stale": true and reports no fraction instead of a number that no longer belongs to the built code.
FLOP_ALGO = {"colav_iw": 370.0 + 26.0, "rl": 450.0 + 31.0}
KERNEL_COUNTS = os.path.join(ROOT, "profiles", "kernel_counts.json")
DEVICE_SOURCES = ("ast_sac_b200/csrc/shipenv_kernels.cuh", "ast_sac_b200/csrc/shipenv_math.cuh",
                  "ast_sac_b200/csrc/kernels_fast.cu", "ast_sac_b200/csrc/Makefile")
# HBM bytes per env-step when every simulator step is its own launch (K = 1): DESIGN.md section 4
BYTES_K1 = 2 * 2 * (17 * 8 + 4) + 2 * (5 * 8 + 2 * 4) + 32 + 8 + 4 + 4     # ship rows r+w, env rows r+w, outputs = 704 B


def device_source_hash() -> str:
    h = hashlib.sha256()
    for rel in DEVICE_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def split_calls(a) -> bool:
    """Does the library issue this workload's step() calls as two k_env launches (csrc/shipenv.cu launch_env)."""
    return (a.workload == "colav_iw" and a.collav == "none" and os.environ.get("SHIPENV_SPLIT_CALLS", "1") != "0"
            and os.environ.get("SHIPENV_QUIET", "1") != "0" and os.environ.get("SHIPENV_PERSISTENT", "1") != "0")


def kernel_counts(workload: str, collav: str):
    """(entry or None, stale flag) of profiles/kernel_counts.json for this workload."""
    try:
        d = json.load(open(KERNEL_COUNTS))
    except Exception:
        return None, True
    key = workload if collav == "none" else f"{workload}+{collav}"
    e = d.get("kernels", {}).get(key)
    return e, (e is None or d.get("source_hash") != device_source_hash())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)     # 20 bench steps = 160 episodes: a timed region of ~1 s at 1e5 envs
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=100_000, help="environments per GPU")
    ap.add_argument("--episodes-per-step", type=int, default=8,
                    help="episodes of every environment per bench step (8 x ~7 ms: K = 20 steps time > 1 s)")
    ap.add_argument("--workload", default="colav_iw", choices=["colav_iw", "rl"])
    ap.add_argument("--collav", default="none", choices=["none", "simple", "sbmpc"],
                    help="collision avoidance of the ship under test (the headline workload uses 'none')")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"], help="device code build (see DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true",
                    help="skip the 'workloads' block (config 3 at 1e6 envs and the SBMPC variants; N = 1 only)")
    ap.add_argument("--no-pin", action="store_true", help="do not pin the rank to its GPU's CPUs")
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


def bench_config(a, world):
    """The `config` object: identical for the b200 and the reference arm (it names the workload, nothing measured)."""
    return {"workload": workload_name(a.workload, a.envs, a.collav), "envs_per_gpu": a.envs, "envs_total": a.envs * world,
            "episodes_per_step": a.episodes_per_step, "collav_mode": a.collav,
            "sharding": "contiguous env blocks per rank, no per-step collective",
            "l2": "256 MB buffer written between timed bench steps (state < 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# rank -> CPU pinning
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_cpus(local_rank: int, local_world: int):
    """Pin this process to a slice of the CPUs that are local to its GPU (NVML affinity mask; falls back to the CPUs
    the process may already run on).  Ranks whose GPUs share a CPU set (one NUMA node for all eight GPUs on this
    pool's boxes) get disjoint slices, so the host side of the e2e path does not migrate or contend."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(local_rank)
            words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            local = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
            local = [c for c in local if c in allowed]
            if local:
                cpus = local
        except Exception:
            pass
        per = max(1, len(cpus) // max(1, local_world))
        mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        return {"cpus": [mine[0], mine[-1]], "n": len(mine)}
    except Exception as exc:
        return {"error": type(exc).__name__}


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
# CPU arm: the oracle port on the host cores, and the unmodified Python reference on one core
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


def reference_python_rate(workload, collav):
    """The UNMODIFIED reference (pure Python, single process) on one host core, from the copy oracle/build_ref.py
    makes (oracle/_ref/, or /root/reference in the build container).  None when neither exists."""
    try:
        from oracle import ref_harness as H
        if not H.reference_available():
            return {"unavailable": f"no reference tree at {H.REFERENCE_ROOT} (oracle/build_ref.py makes oracle/_ref/)"}
        min_steps = 600 if collav == "sbmpc" else 2000
        rate, steps, secs = H.time_reference(workload, collav, min_steps=min_steps, warmup_steps=200)
        return {"value": rate, "unit": METRIC, "cores": 1, "kind": "reference",
                "sample": f"{steps} env-steps of whole step(action) calls after 200 warm-up steps ({secs:.1f} s), the "
                          f"reference's own env class imported from {os.path.relpath(H.REFERENCE_ROOT, ROOT)}",
                "note": "matplotlib / gymnasium / shapely / gtimer are not installed: stub modules of oracle/ref_harness.py "
                        "(shapely: documented-semantics stand-in, pinned by tests/test_map_geometry.py)"}
    except Exception as exc:      # measurement aid: never fail the bench line over it
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def cpu_baseline_block(workload, collav, target_s=12.0):
    n, cores = calibrated_cpu_sample(workload, target_s=target_s, collav=collav)
    rate, steps, secs = cpu_episode_rate(workload, n, collav=collav)
    return {"value": rate, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": f"{n} episodes of the same workload ({steps} env-steps, {secs:.1f} s), "
                      f"oracle/shipsim_oracle.c on {cores} host threads",
            "reference_python_1core": reference_python_rate(workload, collav)}


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
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(a, a.gpus),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference simulator (oracle/), pthreads over episodes; the "
                                 "reference itself is single-process Python, timed on one core beside it",
                         "reference_python_1core": reference_python_rate(a.workload, a.collav)},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """One environment batch on this rank's GPU with its synthetic inputs, and the two ways of running an episode:
    device-resident (torch tensors in, output buffers stay on the device) and through the host-buffer C ABI."""

    def __init__(self, a, workload, collav, B, rank, dev):
        import torch
        from ast_sac_b200 import scenarios as S
        self.torch, self.dev, self.B, self.workload, self.collav = torch, dev, B, workload, collav
        args, assets, m, actions_cpu, init_cpu = make_inputs(workload, B, rank, collav)
        self.actions_dev = actions_cpu.to(dev).t().contiguous()            # [9, B]: one contiguous row per step() call
        self.actions_host = np.ascontiguousarray(actions_cpu.numpy().T)    # [9, B] rows for the host API
        cls = S.MultiShipRLEnv if workload == "rl" else S.MultiShipEnv
        self.env = cls(assets=assets, map=m, args=args, num_envs=B, device=dev, init_states=init_cpu.to(dev),
                       math_mode=a.math)
        self.stream = torch.cuda.current_stream(dev)
        self.ep_return = torch.zeros(B, dtype=torch.float64, device=dev)
        self.ep_rl = torch.zeros(B, dtype=torch.int32, device=dev)
        self.ep_steps = torch.zeros(B, dtype=torch.int32, device=dev)
        self.launch_steps = torch.zeros(1 + N_RL_STEPS, dtype=torch.int64, device=dev)

    def episode(self, events=False, stats=False):
        """reset + 9 step() on device-resident actions; returns the (start, end) event pairs of the 10 calls."""
        torch, env = self.torch, self.env
        if stats:
            self.ep_return.zero_(); self.ep_rl.zero_(); self.ep_steps.zero_(); self.launch_steps.zero_()
        pairs = []
        for j in range(-1, N_RL_STEPS):
            if events:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(self.stream)
            if j < 0:
                env.reset()
            else:
                env.step(self.actions_dev[j])
            if events:
                e.record(self.stream)
                pairs.append((s, e))
            if stats and j >= 0:
                self.ep_return.add_(env.reward_buf)
                self.ep_rl.add_((env.nsub_buf > 0).to(torch.int32))
                self.ep_steps.add_(env.nsub_buf)
                self.launch_steps[j + 1] = env.nsub_buf.sum()
        return pairs

    def host_episode(self):
        """The reference-facing path: numpy actions in, numpy results out, copies inside the call."""
        env = self.env
        env.reset_host()
        for j in range(N_RL_STEPS):
            env.step_host(self.actions_host[j])

    def close(self):
        self.env.close()


def timed_episodes(w, steps, episodes_per_step, l2_flush, barrier):
    """`steps` bench steps of `episodes_per_step` episodes each, timed with CUDA events on the launching stream
    (every reset() / step() call individually, so host gaps between calls are not counted as device time) and, for
    the dominant kernel, with events around k_env itself inside the C ABI.  Returns a dict of raw measurements."""
    import torch
    from ast_sac_b200 import _lib as L
    lib = L.load()
    env = w.env
    c0 = env.total_substeps()
    L.check(lib.shipenv_time_env_kernel(env._handle, 1))
    barrier()
    t_wall0 = time.perf_counter()
    step_ms, launch_ms, kernel_ms = [], [], []
    for _ in range(steps):
        l2_flush.fill_(1.0)                      # > 126 MB L2: evict the state between timed bench steps
        ms_step, k_step = 0.0, 0.0
        for _ in range(episodes_per_step):
            pairs = w.episode(events=True)
            torch.cuda.synchronize(w.dev)
            ms = [s.elapsed_time(e) for s, e in pairs]
            launch_ms.append(ms)
            ms_step += sum(ms)
            kms = ctypes.c_double()
            L.check(lib.shipenv_env_kernel_ms(env._handle, ctypes.byref(kms)))
            k_step += kms.value
        step_ms.append(ms_step)
        kernel_ms.append(k_step)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    L.check(lib.shipenv_time_env_kernel(env._handle, 0))
    return {"steps_done": env.total_substeps() - c0, "time_s": sum(step_ms) / 1e3, "wall_s": t_wall,
            "launch_ms": np.array(launch_ms), "kernel_ms": np.array(kernel_ms), "step_ms": step_ms}


def timed_host_episodes(w, n_episodes, barrier):
    """Episodes through the host-buffer API, wall clock around the whole loop.  The caller's action rows are
    page-locked once (the caller keeps them alive); the executed simulator steps come from the device counter."""
    w.env.register_host_actions(w.actions_host)
    w.host_episode()
    barrier()
    c0 = w.env.total_substeps()
    t0 = time.perf_counter()
    for _ in range(n_episodes):
        w.host_episode()
    barrier()
    t = time.perf_counter() - t0
    n = w.env.total_substeps() - c0
    w.env.unregister_host_actions()
    return n, t


def roofline_block(workload, collav, B, m, n_episodes, fp64_peak):
    """FP64 roofline of the dominant kernel from the measured kernel time and the per-env-step executed flop count of
    profiles/kernel_counts.json (stale: true and no fraction when the device sources changed since the capture)."""
    counts, stale = kernel_counts(workload, collav)
    env_steps = m["steps_done"]
    kernel_s = float(m["kernel_ms"].sum()) * 1e-3
    flop_exec = None if counts is None else counts.get("flop_exec_per_env_step")
    achieved = None if (flop_exec is None or stale) else flop_exec * env_steps / kernel_s / 1e12
    algo = FLOP_ALGO[workload] * env_steps / kernel_s / 1e12 if collav == "none" else None
    return {
        "bound": "fp64", "kernel": "k_env<MODE_STEP> (9 launches per episode)",
        "ms_per_launch": 1e3 * kernel_s / (n_episodes * N_RL_STEPS),
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": None if achieved is None else achieved / fp64_peak, "stale": bool(stale),
        "peak_source": "DFMA microbenchmark measured live in this run (MEASURED_PEAKS.json has no FP64 entry)",
        "flop_per_env_step_executed": flop_exec, "flop_per_env_step_algorithmic": FLOP_ALGO[workload] if collav == "none" else None,
        "achieved_algorithmic": algo,
        "frac_algorithmic": None if algo is None else algo / fp64_peak,
        # ncu's own view of the binding unit in the capture the counts come from: DMUL / DADD / DSETP hold the FP64
        # pipe as long as a DFMA but count one / one / zero flop, so the pipe is busier than `frac` says
        "fp64_pipe_active_pct_ncu": None if (counts is None or stale) else counts.get("pipe_fp64_pct"),
        "issue_slots_active_pct_ncu": None if (counts is None or stale) else counts.get("issue_active_pct"),
        "counts_source": None if counts is None else counts.get("source"),
        # What the executed count covers.  A capture of every launch of an episode counts what the ships really execute:
        # in the last step() call a third of the ship-steps belong to ships that have stopped (they only advance their
        # clock), so the episode's count per env-step is below the count of a call in which both ships sail -- and below
        # SURVEY section 8(d)'s algorithmic figure, which assumes two sailing ships.  Rounds 1 / 2 (first sessions) quoted
        # the fraction with the mid-episode count applied to every env-step: `frac_midcall_count` is that figure.
        "flop_count_scope": None if counts is None else counts.get("scope", "two mid-episode step() calls"),
        "flop_per_env_step_executed_midcall": None if counts is None else counts.get("flop_exec_per_env_step_midcall"),
        "frac_midcall_count": None if (counts is None or stale or not counts.get("flop_exec_per_env_step_midcall")) else
                              counts["flop_exec_per_env_step_midcall"] * env_steps / kernel_s / 1e12 / fp64_peak,
        "traffic": None if (counts is None or stale or B != counts.get("envs")) else counts.get("dram_bytes_per_launch"),
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/)",
        "algorithmic_bytes_per_launch": float(B * (2 * (17 * 8 + 4) * 2 + 2 * (5 * 8 + 8) + 48)),
        "share_of_step": float(m["kernel_ms"].sum() / m["launch_ms"].sum()),
    }


def run_b200(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from ast_sac_b200 import _lib as L
    from ast_sac_b200 import parallel as PAR

    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    pin = None if a.no_pin else pin_to_gpu_cpus(local_rank, local_world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    B, R = a.envs, a.episodes_per_step
    w = Workload(a, a.workload, a.collav, B, rank, dev)
    env = w.env
    fp64_peak = L.measure_fp64_peak(local_rank, repeats=5)
    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    for _ in range(a.warmup):
        w.episode()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = timed_episodes(w, a.steps, R, l2_flush, barrier)
    clocks = sampler.stop()
    n_episodes = a.steps * R
    roofline = roofline_block(a.workload, a.collav, B, m, n_episodes, fp64_peak)

    # ---- HBM roofline of the one-simulator-step-per-launch configuration (K = 1, HBM-bound)
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
        s.record(w.stream)
        # the C ABI entry point itself (env._step() adds torch ops that build the info tensors)
        L.check(L.load().shipenv_substeps(env._handle, 1, env._stream_ptr()))
        e.record(w.stream)
        torch.cuda.synchronize(dev)
        k1_ms.append(s.elapsed_time(e))
    k1 = float(np.mean(k1_ms[4:])) if len(k1_ms) > 8 else float(np.mean(k1_ms))
    k1_gbs = BYTES_K1 * B / (k1 * 1e-3) / 1e9
    roofline_k1 = {"bound": "hbm", "kernel": "k_env<MODE_SUBSTEPS>, k=1 (one _step() per launch)",
                   "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                   "peak_source": hbm_src, "bytes_per_env_step": BYTES_K1, "ms_per_launch": k1,
                   "env_steps_per_s": B / (k1 * 1e-3), "traffic": None}

    # ---- e2e: the reference-facing host API (numpy in / numpy out), H2D + D2H inside the timed region
    if not a.no_e2e:
        n_e2e, t_e2e = timed_host_episodes(w, n_episodes, barrier)
        e2e_local = torch.tensor([n_e2e, t_e2e], dtype=torch.float64, device=dev)
    else:
        e2e_local = torch.zeros(2, dtype=torch.float64, device=dev)

    # ---- aggregate over ranks: sum of steps, max of time; NCCL all-gather of the episode statistics
    agg = torch.tensor([m["steps_done"], m["time_s"], m["wall_s"]], dtype=torch.float64, device=dev)
    w.episode(stats=True)        # a last untimed episode leaves complete per-env statistics in the buffers
    stats = PAR.episode_stats(env.info_buf & L.INFO_EVENT_MASK, w.ep_return, w.ep_steps, w.ep_rl)
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
    launch_steps = [int(x) for x in w.launch_steps.tolist()]
    w.close()

    # ---- the other claimed workloads, same run, same measurement (N = 1 only: the scaling runs stay short)
    extra = None
    if world == 1 and not a.no_extra_workloads:
        extra = {}
        for name, wl, collav, envs in (("rl_1e6 (BASELINE config 3)", "rl", "none", 1_000_000),
                                       ("rl+sbmpc_1e5", "rl", "sbmpc", 100_000),
                                       ("colav_iw+sbmpc_1e5", "colav_iw", "sbmpc", 100_000)):
            x = Workload(a, wl, collav, envs, rank, dev)
            for _ in range(2):
                x.episode()
            n_ep = 4 if envs > 100_000 or collav != "none" else 8
            mm = timed_episodes(x, n_ep, 1, l2_flush, barrier)
            entry = {"workload": workload_name(wl, envs, collav), "value": mm["steps_done"] / mm["time_s"], "unit": METRIC,
                     "episodes_timed": n_ep, "ms_per_episode": 1e3 * mm["time_s"] / n_ep,
                     "kernel_ms_per_episode": float(mm["kernel_ms"].sum() / n_ep),
                     "env_steps_per_episode_mean": mm["steps_done"] / n_ep / envs,
                     "roofline": roofline_block(wl, collav, envs, mm, n_ep, fp64_peak)}
            if not a.no_e2e:
                n_h, t_h = timed_host_episodes(x, max(1, n_ep // 2), barrier)
                entry["e2e"] = {"value": n_h / t_h, "unit": METRIC, "h2d_bytes_per_episode": int(N_RL_STEPS * envs * 8),
                                "d2h_bytes_per_episode": int(N_RL_STEPS * envs * (32 + 8 + 4 + 4) + envs * 32)}
            extra[name] = entry
            x.close()

    if rank == 0:
        launch_ms = m["launch_ms"]
        line = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * max_time / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(a, world),
            "math_mode": a.math,
            "env_steps_per_episode_mean": m["steps_done"] / n_episodes / B,
            "wall_s_timed_region": float(allagg[:, 2].max()), "device_s_timed_region": max_time,
            "clocks": clocks, "cpu_pinning": pin,
            # k_reset + 9 x (k_prologue + k_env) per episode; the colav_iw env without collision avoidance issues every
            # step() call as two k_env launches (quiet-step twin for environments in their last call)
            "gpu_launches": int(n_episodes * (1 + (3 if split_calls(a) else 2) * N_RL_STEPS)),
            "launch_ms_mean": [round(float(x), 4) for x in launch_ms.mean(axis=0)],
            "launch_env_steps": launch_steps,
            "roofline": roofline, "roofline_hbm_k1": roofline_k1,
            "episode_stats": PAR.summarise(gathered),
        }
        if not a.no_e2e:
            e2e_steps, e2e_time = float(alle2e[:, 0].sum()), float(alle2e[:, 1].max())
            line["e2e"] = {"value": e2e_steps / e2e_time, "unit": METRIC,
                           "h2d_bytes_per_step": int(R * N_RL_STEPS * B * 8),
                           "d2h_bytes_per_step": int(R * (N_RL_STEPS * B * (32 + 8 + 4 + 4) + B * 32)),
                           "api": "reset_host() + 9 x step_host(): numpy actions in, numpy obs/reward/info out "
                                  "through shipenv_reset_host / shipenv_step_host (page-locked host arrays)"}
        if extra is not None:
            line["workloads"] = extra
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(a.workload, a.collav)
        print(json.dumps(line), flush=True)
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
