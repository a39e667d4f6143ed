"""One-ulp twins of the detailed-model goldens, run with the UNMODIFIED reference (build container only).

    python tests/golden/make_reference_twins.py            # envelopes of every rl_* golden + the 256-env batches
    python tests/golden/make_reference_twins.py goldens    # only tests/golden/envelope_rl_goldens.npz
    python tests/golden/make_reference_twins.py batch      # only tests/golden/batch_rl_*.npz

Why: with ``measured_shaft_speed = forward_speed`` (rl_env/ship_in_transit/env.py:397-401) the cascaded throttle
controller (rl_env/ship_in_transit/sub_systems/controllers.py:185-189) has gain ~1e4 inside a 1e-4 m/s band, so two
IEEE-correct evaluations of the reference's formulas that differ by one ulp somewhere (another libm, another BLAS)
drift apart by more than 1e-9 for a few percent of the episodes.  How far is measured here ON THE REFERENCE ITSELF:
every episode is re-run twelve times -- four twins with one initial state of both ships moved by one ulp (surge
speed up / down, heading up, shaft speed down), four "other libm" twins in which the reference code runs unchanged
on a math library whose sin / cos / arctan2 / atan results differ from the installed one by at most one ulp (what
any other platform's IEEE-conformant libm does: glibc vs SVML vs CUDA), and four "other BLAS" twins in which
np.dot / np.linalg.inv (the 3x3 products of three_dof_kinetics, rl_env ship_model.py:849-861) return results that
differ by at most one ulp per element (OpenBLAS here vs the MKL of the reference's ast-sac.yml: FMA use and
summation order are the library's choice) -- and the largest relative state difference to the stock run after each
step(action) call is the episode's *reference envelope*.  The GPU tests accept min(1e-6, 10 x envelope) instead of
1e-9 only where the envelope says so, and waive a flag only where a reference twin itself flips it.

Outputs
  envelope_rl_goldens.npz   per rl_* golden: envelope [n_valid] (states), ctrl envelope, reward / obs / travel
                            envelope, and flags_equal [n_valid] (all four twins reproduce the flags of that call)
  batch_rl_<collav>_<B>.npz the reference's own results for the seeded batch of tests/test_gpu_parity.py
                            (test_batched_episodes_match_reference_rl): states, flags, rewards, observations of every
                            environment after every step(action) call, plus the envelopes of its twins
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as H  # noqa: E402
from make_golden import ctrl_vec, events_bits, ship_vec  # noqa: E402

STATE_SCALE = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1e-3, 1.0, 1.0])      # tests/helpers.py
N_FLAGS = 8   # done, events, terminal, test_stop, obs_stop, n_log, k_test, k_obs
OMEGA0 = (420 * np.pi / 30, 200 * np.pi / 30)                            # run/env_setup.py:100,137


def rel_err(a, b, scale):
    return np.abs(a - b) / np.maximum(np.abs(b), scale)


def twin_inits(variant, test_init, obs_init):
    """Initial states of twin `variant` (0 = unperturbed): the same perturbations as
    tests/test_gpu_parity.py:_golden_conditioning applies to the oracle."""
    ti = dict(H.TEST_INIT, **(test_init or {}))
    oi = dict(H.OBS_INIT, **(obs_init or {}))
    om = list(OMEGA0)
    for d in (ti, oi):
        if variant == 1:
            d["initial_forward_speed_m_per_s"] = float(np.nextafter(np.float64(d["initial_forward_speed_m_per_s"]), 10.0))
        elif variant == 2:
            d["initial_forward_speed_m_per_s"] = float(np.nextafter(np.float64(d["initial_forward_speed_m_per_s"]), 0.0))
        elif variant == 3:
            d["initial_yaw_angle_rad"] = float(np.nextafter(np.float64(d["initial_yaw_angle_rad"]), 10.0))
    if variant == 4:
        om = [float(np.nextafter(np.float64(x), 0.0)) for x in om]
    return ti, oi, tuple(om)


N_VARIANTS = 13       # 0 stock, 1-4 one-ulp initial states, 5-8 other-libm twins, 9-12 other-BLAS twins


class _JitterLib:
    """numpy / math with the named transcendental functions moved by -1, 0 or +1 ulp (chosen by a hash of the
    argument bits and the twin's seed): an equally valid (<= 1 ulp) math library beneath the unmodified reference."""

    def __init__(self, base, names, seed):
        self._base, self._seed = base, seed
        for n in names:
            setattr(self, n, self._wrap(getattr(base, n)))

    def __getattr__(self, name):
        return getattr(self._base, name)

    def _wrap(self, fn):
        seed = self._seed

        def jittered(*args):
            r = fn(*args)
            if isinstance(r, np.ndarray):                       # np.dot / np.linalg.inv: every element by -1, 0, +1 ulp
                rng = np.random.default_rng(hash((seed, r.tobytes())) & 0xffffffffffff)
                d = rng.integers(-1, 2, size=r.shape)
                out = np.where(d > 0, np.nextafter(r, np.inf), np.where(d < 0, np.nextafter(r, -np.inf), r))
                return np.where(np.isfinite(r) & (r != 0), out, r)
            h = hash((seed,) + tuple(float(a).hex() for a in args)) % 3
            if h == 0 or not np.isfinite(r) or r == 0:
                return r
            out = np.nextafter(np.float64(r), np.inf if h == 1 else -np.inf)
            return out if isinstance(r, np.floating) else float(out)
        return jittered


def run_episode(job):
    """One episode of the reference's MultiShipRLEnv; returns the per-call results as arrays."""
    variant = job[-1]
    if variant < 5:
        return _run_episode(job)
    import rl_env.ship_in_transit.sub_systems.LOS_guidance as M_los
    import rl_env.ship_in_transit.sub_systems.ship_model as M_ship
    saved = (M_ship.np, M_los.math)
    import math
    if variant < 9:
        M_ship.np = _JitterLib(np, ("sin", "cos", "arctan2"), variant)
        M_los.math = _JitterLib(math, ("atan", "atan2"), variant)
    else:
        blas = _JitterLib(np, ("dot",), variant)
        blas.linalg = _JitterLib(np.linalg, ("inv",), variant)
        M_ship.np = blas
    try:
        return _run_episode(job)
    finally:
        M_ship.np, M_los.math = saved


def _run_episode(job):
    dt, collav, mode, sim_time, test_init, obs_init, actions, variant = job
    ti, oi, om = twin_inits(variant, test_init, obs_init)
    args = H.Args(time_step=dt, collav_mode=collav)
    env, assets = H.make_rl_env(args, mode=mode, test_init=ti, obs_init=oi, sim_time=sim_time, omega_init=om)
    env.reset()
    n = len(actions)
    states = np.zeros((n, 2, 8)); ctrl = np.zeros((n, 2, 7)); obs = np.zeros((n, 8), np.float32)
    reward = np.zeros(n); travel = np.zeros(n); flags = np.zeros((n, N_FLAGS), np.int32)
    n_valid = 0
    for j, a in enumerate(actions):
        o, r, d, info = env.step(np.array([a], dtype=np.float64))
        states[j, 0], states[j, 1] = ship_vec(assets[0]), ship_vec(assets[1])
        ctrl[j, 0], ctrl[j, 1] = ctrl_vec(assets[0]), ctrl_vec(assets[1])
        obs[j] = o
        reward[j] = r
        travel[j] = env.travel_dist
        flags[j] = [bool(d), events_bits(info['events']), bool(info['terminal']), bool(info['test_ship_stop']),
                    bool(info['obs_ship_stop']), len(assets[1].ship_model.simulation_results['time [s]']),
                    assets[0].auto_pilot.next_wpt, assets[1].auto_pilot.next_wpt]
        n_valid = j + 1
        if d:
            break
    return dict(states=states, ctrl=ctrl, obs=obs, reward=reward, travel=travel, flags=flags, n_valid=n_valid)


def envelopes(base, twins):
    """Per step(action) call: the largest relative difference between the unperturbed run and its twins."""
    n = base["n_valid"]
    env_state = np.zeros(n); env_ctrl = np.zeros(n); env_reward = np.zeros(n); env_obs = np.zeros(n)
    env_travel = np.zeros(n); flags_equal = np.ones(n, np.int32)
    for t in twins:
        for j in range(n):
            if j >= t["n_valid"]:
                flags_equal[j] = 0
                continue
            env_state[j] = max(env_state[j], rel_err(t["states"][j], base["states"][j], STATE_SCALE).max())
            env_ctrl[j] = max(env_ctrl[j], rel_err(t["ctrl"][j], base["ctrl"][j], 1.0).max())
            env_reward[j] = max(env_reward[j], float(rel_err(t["reward"][j], base["reward"][j], 1e-3)))
            env_obs[j] = max(env_obs[j], float(np.abs(t["obs"][j].astype(np.float64) - base["obs"][j]).max()))
            env_travel[j] = max(env_travel[j], float(rel_err(t["travel"][j], base["travel"][j], 1.0)))
            if not np.array_equal(t["flags"][j], base["flags"][j]):
                flags_equal[j] = 0
    return dict(state=env_state, ctrl=env_ctrl, reward=env_reward, obs=env_obs, travel=env_travel,
                flags_equal=flags_equal)


def golden_envelopes(pool):
    import glob
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "rl_*.npz")))
    jobs, index = [], []
    for name in names:
        g = np.load(os.path.join(HERE, name + ".npz"))
        meta = json.loads(str(g["meta"]))
        acts = np.asarray(g["actions"], dtype=np.float64)[: int(g["n_valid"])]
        for v in range(N_VARIANTS):
            jobs.append((meta["dt"], meta.get("collav", "none"), meta.get("mode", "PTI"), meta.get("sim_time", 10000),
                         meta.get("test_init"), meta.get("obs_init"), acts, v))
            index.append((name, v))
    t0 = time.time()
    res = pool.map(run_episode, jobs, chunksize=1)
    out = {}
    for name in names:
        runs = [r for (nm, v), r in zip(index, res) if nm == name]
        g = np.load(os.path.join(HERE, name + ".npz"))
        # the unperturbed run must reproduce the committed golden bit for bit
        n = int(g["n_valid"])
        assert runs[0]["n_valid"] == n, name
        assert np.array_equal(runs[0]["states"][:n, 0], g["test_state"][:n]), name
        assert np.array_equal(runs[0]["states"][:n, 1], g["obs_state"][:n]), name
        assert np.array_equal(runs[0]["flags"][:n, 1], g["events"][:n]), name
        e = envelopes(runs[0], runs[1:])
        for k, v in e.items():
            out[f"{name}|{k}"] = v
        print(f"{name:28s} n={n} envelope {np.array2string(e['state'], precision=1)} flags_equal {e['flags_equal']}")
    np.savez_compressed(os.path.join(HERE, "envelope_rl_goldens.npz"), **out)
    print(f"goldens: {len(jobs)} reference episodes in {time.time() - t0:.0f} s")


def batch(pool, collav, B, seed_actions=0, seed_init=1, pos_jitter_m=100.0):
    """The seeded batch of tests/test_gpu_parity.py (same generators, same scaling of the first quarter)."""
    import torch
    from ast_sac_b200 import scenarios as S
    args = S.get_env_args(time_step=4, collav_mode=collav)
    assets, _ = S.build_rl_assets(args)
    init = S.jittered_init_states(assets, B, pos_jitter_m=pos_jitter_m, seed=seed_init, device="cpu")
    init_np = init.numpy().reshape(7, B, 2)
    gen = torch.Generator().manual_seed(seed_actions)
    actions = (torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    actions[: B // 4] *= 0.1
    actions = actions.numpy()
    jobs = []
    for b in range(B):
        ti = dict(initial_north_position_m=float(init_np[0, b, 0]), initial_east_position_m=float(init_np[1, b, 0]))
        oi = dict(initial_north_position_m=float(init_np[0, b, 1]), initial_east_position_m=float(init_np[1, b, 1]))
        for v in range(N_VARIANTS):
            jobs.append((4, collav, "PTI", 10000, ti, oi, actions[b], v))
    t0 = time.time()
    res = pool.map(run_episode, jobs, chunksize=1)
    out = dict(meta=json.dumps(dict(kind="rl", dt=4, collav=collav, B=B, seed_actions=seed_actions, seed_init=seed_init,
                                    pos_jitter_m=pos_jitter_m)),
               actions=actions, init=init_np, states=np.zeros((B, 9, 2, 8)), ctrl=np.zeros((B, 9, 2, 7)),
               obs=np.zeros((B, 9, 8), np.float32), reward=np.zeros((B, 9)), travel=np.zeros((B, 9)),
               flags=np.zeros((B, 9, N_FLAGS), np.int32), n_valid=np.zeros(B, np.int32),
               env_state=np.zeros((B, 9)), env_ctrl=np.zeros((B, 9)), env_reward=np.zeros((B, 9)),
               env_obs=np.zeros((B, 9)), env_travel=np.zeros((B, 9)), flags_equal=np.ones((B, 9), np.int32))
    for b in range(B):
        runs = res[N_VARIANTS * b: N_VARIANTS * (b + 1)]
        base = runs[0]
        n = base["n_valid"]
        for k in ("states", "ctrl", "obs", "reward", "travel", "flags"):
            out[k][b] = base[k]
        out["n_valid"][b] = n
        e = envelopes(base, runs[1:])
        out["env_state"][b, :n] = e["state"]; out["env_ctrl"][b, :n] = e["ctrl"]; out["env_reward"][b, :n] = e["reward"]
        out["env_obs"][b, :n] = e["obs"]; out["env_travel"][b, :n] = e["travel"]; out["flags_equal"][b, :n] = e["flags_equal"]
    path = os.path.join(HERE, f"batch_rl_{collav}_{B}.npz")
    np.savez_compressed(path, **out)
    worst = out["env_state"].max(axis=1)
    print(f"batch {collav} B={B}: {len(jobs)} reference episodes in {time.time() - t0:.0f} s; "
          f"envs with envelope > 1e-9: {(worst > 1e-9).sum()}, > 1e-7: {(worst > 1e-7).sum()}, "
          f"twin flag flips: {(out['flags_equal'] == 0).any(axis=1).sum()}; {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    assert H.reference_available(), "needs /root/reference"
    H.install_stubs()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    with mp.Pool(int(os.environ.get("TWINS_PROCS", os.cpu_count() or 1))) as pool:
        if what in ("all", "goldens"):
            golden_envelopes(pool)
        if what in ("all", "batch"):
            only = sys.argv[2] if len(sys.argv) > 2 else None
            for collav, B in (("none", 256), ("sbmpc", 64), ("simple", 64)):
                if only in (None, collav):
                    batch(pool, collav, B)
