"""Helpers that build product environments from golden-fixture metadata and read their state."""
from __future__ import annotations

import json

import numpy as np

from ast_sac_b200 import scenarios as S
from ast_sac_b200 import _lib as L


def env_from_meta(meta, num_envs=1, init_states=None, device=None, math_mode=None):
    """Build the product env for a golden fixture's ``meta`` (see tests/golden/make_golden.py)."""
    if isinstance(meta, (bytes, str, np.ndarray)):
        meta = json.loads(str(meta))
    args = S.get_env_args(time_step=meta["dt"], collav_mode=meta.get("collav", "none"))
    kw = dict(test_init=meta.get("test_init"), obs_init=meta.get("obs_init"))
    if meta["kind"] == "rl":
        return S.prepare_multiship_rl_env(args, num_envs=num_envs, mode=meta.get("mode", "PTI"), device=device,
                                          init_states=init_states, sim_time=meta.get("sim_time", 10000),
                                          math_mode=math_mode, **kw)
    if meta["kind"] == "colav":
        return S.prepare_colav_env(args, iw=True, num_envs=num_envs, device=device, init_states=init_states,
                                   sim_time=meta.get("sim_time", 10000), math_mode=math_mode, **kw)
    if meta["kind"] == "noniw":
        return S.prepare_colav_env(args, iw=False, num_envs=num_envs, device=device, init_states=init_states,
                                   math_mode=math_mode, **kw)
    if meta["kind"] == "noniw_step":         # MultiShipNonIWEnv driven with step(action)
        return S.prepare_colav_env(args, iw=False, num_envs=num_envs, device=device, init_states=init_states,
                                   sim_time=meta.get("sim_time", 10000), math_mode=math_mode,
                                   obs_route=meta["obs_route"], **kw)
    raise ValueError(meta["kind"])


def assets_from_meta(meta):
    if isinstance(meta, (bytes, str, np.ndarray)):
        meta = json.loads(str(meta))
    args = S.get_env_args(time_step=meta["dt"], collav_mode=meta.get("collav", "none"))
    kw = dict(test_init=meta.get("test_init"), obs_init=meta.get("obs_init"))
    if meta["kind"] == "rl":
        a, m = S.build_rl_assets(args, mode=meta.get("mode", "PTI"), sim_time=meta.get("sim_time", 10000), **kw)
    elif meta["kind"] == "colav":
        a, m = S.build_colav_assets(args, iw=True, sim_time=meta.get("sim_time", 10000), **kw)
    elif meta["kind"] == "noniw_step":
        a, m = S.build_colav_assets(args, iw=False, sim_time=meta.get("sim_time", 10000), obs_route=meta["obs_route"], **kw)
    else:
        a, m = S.build_colav_assets(args, iw=False, **kw)
    return a, m, args


def product_ship_vec(env, role, e=0) -> np.ndarray:
    """[N, E, psi, u, v, r, omega, e_ct] of one ship, in the layout of helpers.STATE_SCALE."""
    s = env.ship_f64[:, 2 * e + role].cpu().numpy()
    has_machinery_state = env._params.ship[0].model_kind != L.MODEL_SIMPLE     # shaft speed or thrust force
    return np.array([s[0], s[1], s[2], s[3], s[4], s[5], s[6] if has_machinery_state else 0.0, s[8]])


def product_ctrl_vec(env, role, e=0) -> np.ndarray:
    """[e_ct_int, hdg_err_i, hdg_prev_err, spd_err_i, spd_prev_err, shaft_err_i, time]."""
    s = env.ship_f64[:, 2 * e + role].cpu().numpy()
    detailed = env._params.ship[0].model_kind == L.MODEL_DETAILED
    sp = [s[12], 0.0, s[13]] if detailed else [s[12], s[13], 0.0]
    return np.array([s[9], s[10], s[11]] + sp + [s[7]])


def product_states_all(env) -> np.ndarray:
    """[num_envs, 2, 8] ship vectors of every environment."""
    s = env.ship_f64.cpu().numpy().reshape(L.SF_COUNT, env.num_envs, 2)
    detailed = env._params.ship[0].model_kind == L.MODEL_DETAILED
    omega = s[6] if detailed else np.zeros_like(s[6])
    return np.stack([s[0], s[1], s[2], s[3], s[4], s[5], omega, s[8]], axis=-1)
