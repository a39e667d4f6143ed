"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference simulator.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is a small .npz holding (a) the oracle config struct bytes extracted from the
reference's own asset objects (so the fixtures are self-contained on the GPU box, where
/root/reference does not exist), (b) the inputs (actions) and (c) the reference outputs.  Map
geometry results come from the Shapely stand-in of oracle/ref_harness.py (Shapely is not installed).
"""
from __future__ import annotations

import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import ref_harness as H  # noqa: E402


def struct_bytes(s) -> np.ndarray:
    return np.frombuffer(ctypes.string_at(ctypes.byref(s), ctypes.sizeof(s)), dtype=np.uint8).copy()


def events_bits(s: str) -> int:
    bits = 0
    rest = s
    for i, ev in enumerate(O.EVENT_STRINGS):
        if ev in rest:
            bits |= 1 << i
            rest = rest.replace(ev, '')
    assert rest == '', f"unparsed events: {rest!r}"
    return bits


def ship_vec(asset):
    sm = asset.ship_model
    omega = sm.ship_machinery_model.omega if hasattr(sm, "ship_machinery_model") else 0.0
    return [float(sm.north), float(sm.east), float(sm.yaw_angle), float(sm.forward_speed),
            float(sm.sideways_speed), float(sm.yaw_rate), float(omega), float(asset.auto_pilot.navigate.e_ct)]


def ctrl_vec(asset):
    ap = asset.auto_pilot
    pid = ap.heading_controller.ship_heading_controller
    if hasattr(asset, "speed_controller"):
        sc = asset.speed_controller.ship_speed_controller
        sp = [float(sc.error_i), float(sc.prev_error), 0.0]
    else:
        tc = asset.throttle_controller
        sp = [float(tc.ship_speed_controller.error_i), 0.0, float(tc.shaft_speed_controller.error_i)]
    return [float(ap.navigate.e_ct_int), float(pid.error_i), float(pid.prev_error)] + sp + [float(asset.ship_model.int.time)]


# ------------------------------------------------------------------------------------------------
def bare_rollout(kind: str, dt, who: int, n_steps: int, mode="PTI", post_reset=False):
    """Bare ship + controllers loop (KAT1 / KAT3 of SURVEY.md section 8c)."""
    args = H.Args(time_step=dt)
    if kind == "simple":
        assets, _ = H.build_colav_assets(args, obs_route="obs_ship_route_nonIW.txt")
    else:
        assets, _ = H.build_rl_assets(args, mode=mode)
    a = assets[who]
    sm = a.ship_model
    if post_reset:
        sm.reset()
        (a.throttle_controller if kind == "detailed" else a.speed_controller).reset()
        a.auto_pilot.reset()
    cfg = O.ship_config_from_asset(a, post_reset=post_reset and kind == "detailed")
    states = np.zeros((n_steps, 8))
    ctrl = np.zeros((n_steps, 7))
    wpt = np.zeros(n_steps, dtype=np.int32)
    for i in range(n_steps):
        rud = a.auto_pilot.rudder_angle_from_sampled_route(sm.north, sm.east, sm.yaw_angle)
        if kind == "simple":
            cmd = a.speed_controller.thrust(a.desired_forward_speed, sm.forward_speed)
            sm.update_differentials(thrust_force=cmd, rudder_angle=rud)
        else:
            cmd = a.throttle_controller.throttle(speed_set_point=a.desired_forward_speed,
                                                 measured_speed=sm.forward_speed,
                                                 measured_shaft_speed=sm.forward_speed)
            sm.update_differentials(engine_throttle=cmd, rudder_angle=rud)
        sm.integrate_differentials()
        sm.int.next_time()
        states[i] = ship_vec(a)
        ctrl[i] = ctrl_vec(a)
        wpt[i] = a.auto_pilot.next_wpt
    # keep the first 256 steps, then every 16th, and always the last
    keep = np.unique(np.concatenate([np.arange(min(256, n_steps)), np.arange(15, n_steps, 16), [n_steps - 1]]))
    meta = dict(kind=kind, dt=dt, who=who, mode=mode, post_reset=post_reset)
    return dict(cfg=struct_bytes(cfg), n_steps=n_steps, step_index=keep + 1, states=states[keep],
                ctrl=ctrl[keep], next_wpt=wpt, dt_shaft=cfg.dt_shaft, meta=json.dumps(meta))


def iw_episode(kind: str, dt, actions, collav="none", mode="PTI", test_init=None, obs_init=None, sim_time=10000,
               obs_route="obs_ship_route.txt"):
    """Episode of run_colav.MultiShipEnv ("colav"), rl_env MultiShipRLEnv ("rl"): KAT2 / KAT4, or of
    run_colav.MultiShipNonIWEnv.step(action) ("noniw_step", run_colav/env.py:678-800: the NonIW class stepped with
    scoping angles -- its obstacle ship must then carry a HeadingBySampledRouteController)."""
    args = H.Args(time_step=dt, collav_mode=collav)
    if kind == "colav":
        env, assets = H.make_colav_iw_env(args, test_init=test_init, obs_init=obs_init, sim_time=sim_time)
        env_kind = O.ENV_COLAV_IW
    elif kind == "noniw_step":
        env, assets = H.make_colav_noniw_env(args, obs_route=obs_route, test_init=test_init, obs_init=obs_init,
                                             sim_time=sim_time)
        env_kind = O.ENV_COLAV_NONIW
    else:
        env, assets = H.make_rl_env(args, mode=mode, test_init=test_init, obs_init=obs_init, sim_time=sim_time)
        env_kind = O.ENV_RL
    cfg = O.env_config_from_assets(assets, env.map, args, env_kind)
    obs0 = env.reset()
    n = len(actions)
    meta = dict(kind=kind, dt=dt, collav=collav, mode=mode, test_init=test_init, obs_init=obs_init, sim_time=sim_time)
    if kind == "noniw_step":
        meta["obs_route"] = obs_route
    out = dict(meta=json.dumps(meta), cfg=struct_bytes(cfg), actions=np.asarray(actions, dtype=np.float64), obs0=np.asarray(obs0),
               obs=np.zeros((n, 8), np.float32), reward=np.zeros(n), done=np.zeros(n, np.int32),
               events=np.zeros(n, np.int32), terminal=np.zeros(n, np.int32), test_stop=np.zeros(n, np.int32),
               obs_stop=np.zeros(n, np.int32), n_log=np.zeros(n, np.int32), k_test=np.zeros(n, np.int32),
               k_obs=np.zeros(n, np.int32), test_state=np.zeros((n, 8)), obs_state=np.zeros((n, 8)),
               test_ctrl=np.zeros((n, 7)), obs_ctrl=np.zeros((n, 7)), travel_dist=np.zeros(n),
               n_valid=0)
    for j, a in enumerate(actions):
        res = env.step(np.array([a], dtype=np.float64))
        if kind in ("colav", "noniw_step"):
            o, d, info = res
            r = 0.0
        else:
            o, r, d, info = res
        out["obs"][j, :len(o)] = o        # (six entries in the NonIW env)
        out["reward"][j] = r
        out["done"][j] = bool(d)
        out["events"][j] = events_bits(info['events'])
        out["terminal"][j] = bool(info['terminal'])
        out["test_stop"][j] = bool(info['test_ship_stop'])
        out["obs_stop"][j] = bool(info['obs_ship_stop'])
        out["n_log"][j] = len(assets[1].ship_model.simulation_results['time [s]'])
        out["k_test"][j] = assets[0].auto_pilot.next_wpt
        out["k_obs"][j] = assets[1].auto_pilot.next_wpt
        out["test_state"][j] = ship_vec(assets[0])
        out["obs_state"][j] = ship_vec(assets[1])
        out["test_ctrl"][j] = ctrl_vec(assets[0])
        out["obs_ctrl"][j] = ctrl_vec(assets[1])
        out["travel_dist"][j] = env.travel_dist
        out["n_valid"] = j + 1
        if d:
            break
    out["obs_route_north"] = np.asarray(assets[1].auto_pilot.navigate.north, dtype=np.float64)
    out["obs_route_east"] = np.asarray(assets[1].auto_pilot.navigate.east, dtype=np.float64)
    if kind == "rl":
        out["substep_rewards"] = np.asarray(env.reward_tracker.total, dtype=np.float64)
    # per-substep trajectory of both ships from the reference's own log (pre-integration rows)
    for who, name in ((0, "test"), (1, "obs")):
        sr = assets[who].ship_model.simulation_results
        out[f"{name}_log_north"] = np.asarray(sr['north position [m]'], dtype=np.float64)
        out[f"{name}_log_east"] = np.asarray(sr['east position [m]'], dtype=np.float64)
        out[f"{name}_log_ect"] = np.asarray(sr['cross track error [m]'], dtype=np.float64)
    return out


def noniw_run(dt, collav="none", max_steps=4000, test_init=None, obs_init=None, use_reset=False):
    """run_colav.MultiShipNonIWEnv: init_step() then _step() until combined_done (config 1)."""
    args = H.Args(time_step=dt, collav_mode=collav)
    env, assets = H.make_colav_noniw_env(args, test_init=test_init, obs_init=obs_init)
    cfg = O.env_config_from_assets(assets, env.map, args, O.ENV_COLAV_NONIW)
    if use_reset:
        env.reset()
    else:
        env.init_step()
    obs, ev, term, ts, os_, done, tst, ost, kt, ko = [], [], [], [], [], [], [], [], [], []
    for _ in range(max_steps):
        o, d, info = env._step()
        obs.append(o); ev.append(events_bits(info['events'])); term.append(bool(info['terminal']))
        ts.append(bool(info['test_ship_stop'])); os_.append(bool(info['obs_ship_stop'])); done.append(bool(d))
        tst.append(ship_vec(assets[0]) + [assets[0].ship_model.int.time])
        ost.append(ship_vec(assets[1]) + [assets[1].ship_model.int.time])
        kt.append(assets[0].auto_pilot.next_wpt); ko.append(assets[1].auto_pilot.next_wpt)
        if d:
            break
    meta = dict(kind="noniw", dt=dt, collav=collav, test_init=test_init, obs_init=obs_init, use_reset=use_reset)
    return dict(meta=json.dumps(meta), cfg=struct_bytes(cfg), obs=np.asarray(obs, np.float32), events=np.asarray(ev, np.int32),
                terminal=np.asarray(term, np.int32), test_stop=np.asarray(ts, np.int32),
                obs_stop=np.asarray(os_, np.int32), done=np.asarray(done, np.int32),
                test_state=np.asarray(tst), obs_state=np.asarray(ost), k_test=np.asarray(kt, np.int32),
                k_obs=np.asarray(ko, np.int32), use_reset=int(use_reset))


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name:40s} {os.path.getsize(path) / 1024:8.1f} KiB")


def main():
    assert H.reference_available(), "needs /root/reference"
    deg = np.deg2rad
    kat_actions = deg(np.array([-2, 0, 5, -5, 10, 0, 0, 0, 0], dtype=np.float64))
    rng = np.random.default_rng(20261018)

    # --- bare ship loops (KAT1, KAT3) ---
    for dt in (1, 4, 30):
        for who in (0, 1):
            save(f"bare_simple_dt{dt}_{'test' if who == 0 else 'obs'}", bare_rollout("simple", dt, who, 10000))
    for mode in ("PTI", "PTO", "MEC"):
        save(f"bare_detailed_{mode}_dt4_prereset", bare_rollout("detailed", 4, 0, 4000, mode=mode))
        save(f"bare_detailed_{mode}_dt4_postreset", bare_rollout("detailed", 4, 0, 4000, mode=mode, post_reset=True))
    save("bare_detailed_PTI_dt4_obs_postreset", bare_rollout("detailed", 4, 1, 4000, post_reset=True))

    # --- IW episodes (KAT2, KAT4) ---
    save("colav_iw_dt4_kat", iw_episode("colav", 4, kat_actions))
    save("rl_dt4_kat", iw_episode("rl", 4, kat_actions))
    for i in range(4):
        acts = rng.uniform(-np.pi / 6, np.pi / 6, size=9)
        save(f"colav_iw_dt4_rand{i}", iw_episode("colav", 4, acts))
        save(f"rl_dt4_rand{i}", iw_episode("rl", 4, acts))
    # large scoping angles: waypoints land on islands / outside the map -> sampling failure path
    save("colav_iw_dt4_fail", iw_episode("colav", 4, deg(np.array([10., 30., 30., 30., 30., 30., 30., 30., 30.]))))
    save("rl_dt4_fail", iw_episode("rl", 4, deg(np.array([10., -30., -30., -30., -30., -30., -30., -30., -30.]))))
    # small scoping angles keep the obstacle ship on the test ship's track -> "Ships collision!"
    # (action rows found by scanning 20000 random episodes with the oracle)
    save("rl_dt4_collision", iw_episode("rl", 4, np.array(
        [-0.04947389, 0.02654783, 0.00399436, -0.01783045, 0.03020418, -0.02060939, -0.00486969, -0.03832306,
         -0.01014598])))
    save("colav_iw_dt4_collision", iw_episode("colav", 4, np.array(
        [0.02003677, -0.03365987, -0.01086403, -0.05174993, -0.0248715, -0.00825309, -0.04126783, 0.01394448,
         -0.01252194])))
    # obstacle ship navigational failure (cross-track error > 500 m)
    nav = np.array([-0.07864699, -0.28520827, 0.10994892, -0.21765305, -0.17804338, 0.00772443, -0.18959464,
                    -0.25623406, -0.51439543])
    save("rl_dt4_obsnavfail", iw_episode("rl", 4, nav))
    save("colav_iw_dt4_obsnavfail", iw_episode("colav", 4, nav))
    # test ship starts next to island 1 and runs aground within a few steps
    aground = dict(test_init=dict(initial_north_position_m=2380.0, initial_east_position_m=200.0))
    save("rl_dt4_testgrounding", iw_episode("rl", 4, kat_actions, **aground))
    save("colav_iw_dt4_testgrounding", iw_episode("colav", 4, kat_actions, **aground))
    # simulation time limit (sim_time = 1000 s -> 250 steps)
    save("rl_dt4_timelimit", iw_episode("rl", 4, kat_actions, sim_time=1000))
    save("colav_iw_dt4_timelimit", iw_episode("colav", 4, kat_actions, sim_time=1000))
    save("rl_dt4_simple_collav", iw_episode("rl", 4, kat_actions, collav="simple"))
    save("colav_iw_dt4_simple_collav", iw_episode("colav", 4, kat_actions, collav="simple"))
    save("rl_dt10_kat", iw_episode("rl", 10, kat_actions))
    save("rl_dt4_PTO", iw_episode("rl", 4, kat_actions, mode="PTO"))
    save("rl_dt4_MEC", iw_episode("rl", 4, kat_actions, mode="MEC"))

    # --- SBMPC collision avoidance (the trainer's default collav mode) ---
    save("rl_dt4_sbmpc_kat", iw_episode("rl", 4, kat_actions, collav="sbmpc"))
    save("colav_iw_dt4_sbmpc_kat", iw_episode("colav", 4, kat_actions, collav="sbmpc"))
    # small scoping angles: the ships meet head-on on the shared track, SBMPC active for a long stretch
    save("rl_dt4_sbmpc_encounter", iw_episode("rl", 4, np.array(
        [-0.04947389, 0.02654783, 0.00399436, -0.01783045, 0.03020418, -0.02060939, -0.00486969, -0.03832306,
         -0.01014598]), collav="sbmpc"))
    save("colav_iw_dt4_sbmpc_encounter", iw_episode("colav", 4, np.array(
        [0.02003677, -0.03365987, -0.01086403, -0.05174993, -0.0248715, -0.00825309, -0.04126783, 0.01394448,
         -0.01252194]), collav="sbmpc"))
    save("rl_dt4_sbmpc_rand0", iw_episode("rl", 4, np.random.default_rng(777).uniform(-np.pi / 6, np.pi / 6, size=9) * 0.3,
                                          collav="sbmpc"))
    save("colav_noniw_dt4_sbmpc", noniw_run(4, collav="sbmpc"))

    # --- NonIW env (config 1) ---
    save("colav_noniw_dt4", noniw_run(4))
    save("colav_noniw_dt30", noniw_run(30))
    save("colav_noniw_dt4_simple_collav", noniw_run(4, collav="simple"))
    save("colav_noniw_dt4_reset", noniw_run(4, use_reset=True))


# ------------------------------------------------------------------------------------------------
# sampler goldens: the reference's own ast_sac_rollout through its NormalizedBoxEnv
# ------------------------------------------------------------------------------------------------
SAMPLER_POLICIES = {   # a = tanh(w3 * obs[3] + w4 * obs[4] + w5 * obs[5] + b), evaluated in float32
    "sampler_rl_dt4_pol0": (0.00012, -0.00009, 0.3, 0.35),
    "sampler_rl_dt4_pol1": (0.00002, -0.00003, 0.05, 0.25),
    "sampler_rl_dt4_pol2": (-0.00001, 0.00002, -0.02, -0.1),
    "sampler_rl_dt4_sbmpc_pol1": (0.00002, -0.00003, 0.05, 0.25),
}


class LinearTanhPolicy:
    """Deterministic stand-in for the TanhGaussian policy: float32 arithmetic, action in [-1, 1]."""

    def __init__(self, w):
        self.w = [np.float32(x) for x in w]

    def reset(self):
        pass

    def get_action(self, o):
        o = np.asarray(o, dtype=np.float32)
        z = self.w[0] * o[3] + self.w[1] * o[4] + self.w[2] * o[5] + self.w[3]
        return np.array([np.tanh(z)], dtype=np.float32), {}


def sampler_golden(w, collav="none", reward_scale=0.75, max_path_length=9):
    from ast_sac.samplers.data_collector.rollout_functions import ast_sac_rollout
    from ast_sac.env_wrapper.normalized_box_env import NormalizedBoxEnv
    import copy
    env, _ = H.make_rl_env(H.Args(time_step=4, collav_mode=collav))

    class Recording(NormalizedBoxEnv):
        """Keeps a copy of every env_info at the time step() returned it: the reference appends the env's own
        snapshot dict to the path, and the sampling-failure branch (rl_env env.py:673-693) later mutates that
        same object, so path['env_infos'][-2] is retroactively overwritten."""
        infos = []

        def step(self, action):
            out = super().step(action)
            self.infos.append(copy.deepcopy(out[3]))
            return out

    wrapped = Recording(env, reward_scale=reward_scale)
    wrapped.infos = []
    path = ast_sac_rollout(wrapped, LinearTanhPolicy(w), max_path_length=max_path_length)
    path["env_infos"] = wrapped.infos
    meta = dict(kind="rl", dt=4, collav=collav, reward_scale=reward_scale, max_path_length=max_path_length, w=list(w))
    return dict(meta=json.dumps(meta), observations=np.asarray(path["observations"], dtype=np.float32),
                actions=np.asarray(path["actions"], dtype=np.float32), rewards=np.asarray(path["rewards"], dtype=np.float64),
                next_observations=np.asarray(path["next_observations"], dtype=np.float32),
                terminals=np.asarray(path["terminals"]).astype(np.uint8), dones=np.asarray(path["dones"]).astype(np.uint8),
                events=np.array([events_bits(i["events"]) for i in path["env_infos"]], dtype=np.int32))


# ------------------------------------------------------------------------------------------------
# trajectory-log goldens: the reference's simulation_results of one KAT episode
# ------------------------------------------------------------------------------------------------
LOG_KEYS = ['time [s]', 'north position [m]', 'east position [m]', 'yaw angle [deg]', 'rudder angle [deg]',
            'forward speed [m/s]', 'sideways speed [m/s]', 'yaw rate [deg/sec]', 'thrust force [kN]',
            'cross track error [m]', 'heading error [deg]', 'propeller shaft speed [rpm]']


def log_golden(kind):
    args = H.Args(time_step=4)
    env, assets = (H.make_colav_iw_env(args) if kind == "colav" else H.make_rl_env(args))
    env.reset()
    actions = np.deg2rad(np.array([-2, 0, 5, -5, 10, 0, 0, 0, 0], dtype=np.float64))
    for a in actions:
        res = env.step(np.array([a], dtype=np.float64))
        if res[-2]:
            break
    out = dict(meta=json.dumps(dict(kind=kind, dt=4, collav="none")), actions=actions)
    for who, name in ((0, "test"), (1, "obs")):
        sr = assets[who].ship_model.simulation_results
        n = len(sr['time [s]'])
        rows = np.unique(np.concatenate([np.arange(0, n, 7), np.arange(max(0, n - 60), n)]))
        out[f"{name}_n"] = n
        out[f"{name}_rows"] = rows
        for k in sr:        # every key of simulation_results: 12 state / controller columns + the machinery bookkeeping
            if len(sr[k]) == n:
                out[f"{name}|{k}"] = np.asarray(sr[k], dtype=np.float64)[rows]
    return out


# ------------------------------------------------------------------------------------------------
# A8' of SURVEY.md section 8a: SimplifiedMachineryModel (thrust-force state T).  No ship model class of the
# reference consumes it, so the golden composes the reference's OWN pieces the way ShipModelAST composes the
# detailed machinery (rl_env ship_model.py:882-901): the hull of a ShipModelAST instance whose
# ship_machinery_model is replaced by an adapter around the reference's SimplifiedMachineryModel, driven by the
# reference's ThrottleFromSpeedSetPointSimplifiedPropulsion and HeadingBySampledRouteController.
# ------------------------------------------------------------------------------------------------
SIMPLIFIED = dict(thrust_force_dynamic_time_constant=30.0, initial_thrust_force=0.0, kp=3.0, ki=0.02)


def bare_simplified(dt, n_steps):
    from rl_env.ship_in_transit.sub_systems.ship_engine import (SimplifiedMachineryModel,
                                                               SimplifiedPropulsionMachinerySystemConfiguration)
    from rl_env.ship_in_transit.sub_systems.controllers import ThrottleFromSpeedSetPointSimplifiedPropulsion
    args = H.Args(time_step=dt)
    assets, _ = H.build_rl_assets(args, mode="PTI")
    a = assets[0]
    sm = a.ship_model
    full = sm.ship_machinery_model
    cfg_m = SimplifiedPropulsionMachinerySystemConfiguration(
        hotel_load=full.hotel_load, machinery_modes=full.machinery_modes, machinery_operating_mode=0,
        specific_fuel_consumption_coefficients_me=full.fuel_coeffs_for_main_engine,
        specific_fuel_consumption_coefficients_dg=full.fuel_coeffs_for_diesel_gen,
        thrust_force_dynamic_time_constant=SIMPLIFIED["thrust_force_dynamic_time_constant"],
        rudder_angle_to_sway_force_coefficient=full.c_rudder_v, rudder_angle_to_yaw_force_coefficient=full.c_rudder_r,
        max_rudder_angle_degrees=30)
    simp = SimplifiedMachineryModel(cfg_m, time_step=dt, initial_thrust_force=SIMPLIFIED["initial_thrust_force"])

    class Adapter:                      # the three calls ShipModelAST makes on its machinery (ship_model.py:882-901)
        c_rudder_v, c_rudder_r = simp.c_rudder_v, simp.c_rudder_r

        def update_shaft_equation(self, load_perc):
            simp.update_thrust_force(load_perc)

        def thrust(self):
            return simp.thrust

        def integrate_differentials(self):
            simp.integrate_differentials()

    sm.ship_machinery_model = Adapter()
    ctrl_ = ThrottleFromSpeedSetPointSimplifiedPropulsion(kp=SIMPLIFIED["kp"], ki=SIMPLIFIED["ki"], time_step=dt)
    cfg = O.ship_config_from_asset(assets[0].__class__(**{**assets[0].__dict__, "ship_model": _with_machinery(sm, full)}))
    cfg.model_kind = O.MODEL_SIMPLIFIED
    cfg.kp_ship_speed, cfg.ki_ship_speed, cfg.dt_shaft = SIMPLIFIED["kp"], SIMPLIFIED["ki"], float(dt)
    states = np.zeros((n_steps, 8))
    wpt = np.zeros(n_steps, dtype=np.int32)
    for i in range(n_steps):
        rud = a.auto_pilot.rudder_angle_from_sampled_route(sm.north, sm.east, sm.yaw_angle)
        cmd = ctrl_.throttle(speed_set_point=a.desired_forward_speed, measured_speed=sm.forward_speed)
        sm.update_differentials(engine_throttle=cmd, rudder_angle=rud)
        sm.integrate_differentials()
        sm.int.next_time()
        states[i] = [float(sm.north), float(sm.east), float(sm.yaw_angle), float(sm.forward_speed),
                     float(sm.sideways_speed), float(sm.yaw_rate), float(simp.thrust), float(a.auto_pilot.navigate.e_ct)]
        wpt[i] = a.auto_pilot.next_wpt
    keep = np.unique(np.concatenate([np.arange(min(256, n_steps)), np.arange(15, n_steps, 16), [n_steps - 1]]))
    meta = dict(kind="simplified", dt=dt, who=0, mode="PTI", **SIMPLIFIED)
    return dict(cfg=struct_bytes(cfg), n_steps=n_steps, step_index=keep + 1, states=states[keep], next_wpt=wpt,
                err_i=float(ctrl_.ship_speed_controller.error_i), meta=json.dumps(meta))


def _with_machinery(sm, full):
    """A shallow view of the ship model that still exposes the detailed machinery object, for config extraction."""
    import copy
    v = copy.copy(sm)
    v.ship_machinery_model = full
    return v


def main_simplified():
    H.install_stubs()
    out = bare_simplified(4, 4000)
    np.savez_compressed(os.path.join(HERE, "bare_simplified_dt4_test.npz"), **out)
    print("simplified", out["states"][-1])


def main_logs():
    H.install_stubs()
    for kind in ("colav", "rl"):
        out = log_golden(kind)
        np.savez_compressed(os.path.join(HERE, f"log_{kind}_dt4_kat.npz"), **out)
        print("log", kind, out["test_n"], out["obs_n"])


def main_sampler():
    H.install_stubs()
    for name, w in SAMPLER_POLICIES.items():
        out = sampler_golden(w, collav="sbmpc" if "sbmpc" in name else "none")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, len(out["actions"]), out["actions"].ravel(), out["events"][-1])


def main_noniw_step():
    """MultiShipNonIWEnv.step(action) (run_colav/env.py:678-800): the NonIW class driven with scoping angles.  Its
    obs_step runs the collision avoidance for the obstacle ship too, has no travel tracker, and the observation has
    six entries."""
    deg = np.deg2rad
    kat_actions = deg(np.array([-2, 0, 5, -5, 10, 0, 0, 0, 0], dtype=np.float64))
    rng = np.random.default_rng(20261019)
    save("colav_stepniw_dt4_kat", iw_episode("noniw_step", 4, kat_actions))
    save("colav_stepniw_dt4_rand0", iw_episode("noniw_step", 4, rng.uniform(-np.pi / 6, np.pi / 6, size=9)))
    save("colav_stepniw_dt4_fail", iw_episode("noniw_step", 4, deg(np.array([10., 30., 30., 30., 30., 30., 30., 30., 30.]))))
    enc = np.array([0.02003677, -0.03365987, -0.01086403, -0.05174993, -0.0248715, -0.00825309, -0.04126783,
                    0.01394448, -0.01252194])
    save("colav_stepniw_dt4_collision", iw_episode("noniw_step", 4, enc))
    save("colav_stepniw_dt4_sbmpc_encounter", iw_episode("noniw_step", 4, enc, collav="sbmpc"))
    save("colav_stepniw_dt4_simple_collav", iw_episode("noniw_step", 4, kat_actions, collav="simple"))
    save("colav_stepniw_dt4_timelimit", iw_episode("noniw_step", 4, kat_actions, sim_time=1000))
    # the 11-waypoint file route of run_simplified_model.py: intermediate waypoints are inserted before its last point
    save("colav_stepniw_dt4_longroute", iw_episode("noniw_step", 4, kat_actions * 0.5,
                                                   obs_route="obs_ship_route_nonIW.txt"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "noniw_step":
        main_noniw_step()
    elif len(sys.argv) > 1 and sys.argv[1] == "sampler":
        main_sampler()      # only the sampler fixtures (the others are unchanged)
    elif len(sys.argv) > 1 and sys.argv[1] == "logs":
        main_logs()
    elif len(sys.argv) > 1 and sys.argv[1] == "simplified":
        main_simplified()
    else:
        main()
        main_sampler()
        main_logs()
        main_simplified()
        main_noniw_step()
