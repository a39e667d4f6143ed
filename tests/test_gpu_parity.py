"""GPU parity tests (run with -m gpu on the B200 box).  Every test drives the CUDA kernels through
the C ABI (include/shipenv.h via ast_sac_b200) and checks them against

  * the golden vectors produced by the unmodified Python reference (tests/golden/*.npz), and
  * the CPU oracle (oracle/shipsim_oracle.c) on the same seeded inputs,

with flags / event bits / waypoint indices / step counts bit-exact and FP64 states within
REL_TOL = 1e-9 relative (BASELINE.json north_star), plus size-independent properties at
BASELINE.json's full sizes (1e5 / 1e6 environments).
"""
import ctypes as C
import json

import numpy as np
import pytest
import torch

from ast_sac_b200 import _lib as L
from ast_sac_b200 import scenarios as S
from oracle import oracle as O

from helpers import (CTRL_SCALE, REFERENCE_BATCHES, REL_TOL, STATE_SCALE, compare_with_reference_batch, envelope_tol, golden,
                     golden_envelope, golden_names, oracle_ctrl_vec, oracle_ship_vec, rel_err, struct_from_bytes)
from product_helpers import assets_from_meta, env_from_meta, product_ctrl_vec, product_ship_vec, product_states_all

pytestmark = pytest.mark.gpu

MATH_MODES = ["strict", "fast"]     # both builds of the device code are held to the same bar


def _sync():
    torch.cuda.synchronize()


def test_device_math_equals_cuda_math_library_bit_for_bit():
    """csrc/shipenv_math.cuh: constant-bank sincos / atan / exp / atan2, the loop-free fmod (both builds) and the
    fast build's branch-free sqrt / division return the CUDA math library's bits (2^24 + 2^22 pseudo-random
    arguments, two seeds)."""
    for n, seed in ((1 << 24, 1), (1 << 22, 12345)):
        counts = L.selftest_math(0, n, seed)
        assert counts == [0] * 18, f"sincos/atan/sqrt/div/exp/atan2/fmod/rsqrt forms mismatches (fast, strict) = {counts}"


# ------------------------------------------------------------------------------------------------
# golden fixtures from the reference: IW episodes (KAT2 / KAT4 and the rare-event cases)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("math_mode", MATH_MODES)
@pytest.mark.parametrize("name", golden_names("colav_iw_") + golden_names("rl_") + golden_names("colav_stepniw_"))
def test_iw_episode_matches_reference_golden(name, math_mode):
    g = golden(name)
    meta = json.loads(str(g["meta"]))
    env, assets = env_from_meta(meta, math_mode=math_mode)
    is_rl = meta["kind"] == "rl"
    detailed = is_rl
    obs0 = env.reset()
    assert np.array_equal(np.asarray(obs0), g["obs0"])
    n = int(g["n_valid"])
    # Detailed model: the bar is 1e-9, widened only where the UNMODIFIED reference's own one-ulp twins of this
    # episode drift apart (tests/golden/envelope_rl_goldens.npz, made by make_reference_twins.py); flags are
    # bit-exact without exception (no reference twin flips one in any golden).
    ref_env = golden_envelope(name) if detailed else None
    widened = []
    for j in range(n):
        res = env.step(np.array([g["actions"][j]]))
        if is_rl:
            o, r, d, info = res
        else:
            o, d, info = res
            r = 0.0
        bits = int(env.info_buf[0].item())
        assert d == bool(g["done"][j]), (name, j)
        assert (bits & L.INFO_EVENT_MASK) == g["events"][j], (name, j, info["events"])
        assert info["terminal"] == bool(g["terminal"][j])
        assert info["test_ship_stop"] == bool(g["test_stop"][j])
        assert info["obs_ship_stop"] == bool(g["obs_stop"][j])
        k = env.next_wpt[0].cpu().numpy()
        assert k[0] == g["k_test"][j] and k[1] == g["k_obs"][j], (name, j, k)
        tol = envelope_tol(ref_env["state"][j]) if ref_env else REL_TOL
        ctrl_tol = envelope_tol(max(ref_env["state"][j], ref_env["ctrl"][j])) if ref_env else REL_TOL
        loose = tol / REL_TOL
        for role, key in ((0, "test"), (1, "obs")):
            e = rel_err(product_ship_vec(env, role), g[key + "_state"][j], STATE_SCALE)
            assert e.max() < tol, (name, j, key, e, tol)
            if e.max() >= REL_TOL:
                widened.append((j, key, float(e.max()), float(ref_env["state"][j])))
            e = rel_err(product_ctrl_vec(env, role), g[key + "_ctrl"][j], CTRL_SCALE)
            assert e.max() < ctrl_tol, (name, j, key, "ctrl", e, ctrl_tol)
        assert rel_err(float(env.env_f64[L.EF["travel_dist"], 0]), g["travel_dist"][j], 1.0) < tol
        # (six observation entries in MultiShipNonIWEnv, colav_stepniw_*; the fixture rows are zero-padded to eight)
        np.testing.assert_allclose(o, g["obs"][j][:len(o)], rtol=2e-7 * loose, atol=1e-6 * loose)
        assert len(o) == (6 if meta["kind"] == "noniw_step" else 8)
        if is_rl:
            assert rel_err(r, g["reward"][j], 1e-3) < 1e-8 * loose, (name, j, r, g["reward"][j])
    if widened:
        print(f"[{name}/{math_mode}] calls beyond 1e-9 but inside the reference's one-ulp envelope (call, ship, err, "
              f"reference envelope): {widened}")
    # number of simulator steps: the reference log has one row per _step() plus the init_step row
    # (a sampling failure adds none)
    assert env.total_substeps() == int(g["n_log"][n - 1]) - 1
    rn, re = env.obs_route()
    assert rel_err(rn, g["obs_route_north"], 1.0).max() < 1e-12
    assert rel_err(re, g["obs_route_east"], 1.0).max() < 1e-12
    env.close()


@pytest.mark.parametrize("math_mode", MATH_MODES)
@pytest.mark.parametrize("name", golden_names("colav_noniw_"))
def test_noniw_run_matches_reference_golden(name, math_mode):
    """config 1 (run_colav/run_simplified_model.py): init_step() + _step() loop, one launch per step."""
    g = golden(name)
    meta = json.loads(str(g["meta"]))
    env, _ = env_from_meta(meta, math_mode=math_mode)
    if meta["use_reset"]:
        env.reset()
    else:
        env.init_step()
    n = len(g["done"])
    tol = 1e-6 if meta["dt"] == 30 else REL_TOL   # dt = 30: explicit-Euler amplification, SURVEY.md section 7
    for i in range(n):
        o, d, info = env._step()
        bits = int(env.info_buf[0].item())
        assert d == bool(g["done"][i]), (name, i)
        assert (bits & L.INFO_EVENT_MASK) == g["events"][i], (name, i, info["events"])
        assert info["terminal"] == bool(g["terminal"][i]) and info["test_ship_stop"] == bool(g["test_stop"][i])
        assert info["obs_ship_stop"] == bool(g["obs_stop"][i])
        k = env.next_wpt[0].cpu().numpy()
        assert k[0] == g["k_test"][i] and k[1] == g["k_obs"][i]
        if i % 16 == 0 or i > n - 40:
            for role, key in ((0, "test_state"), (1, "obs_state")):
                v = np.append(product_ship_vec(env, role), env.read_ship_state(role)[L.SF["time"]])
                e = rel_err(v, g[key][i], np.append(STATE_SCALE, 1.0))
                assert e.max() < tol, (name, i, key, e)
        np.testing.assert_allclose(o, g["obs"][i], rtol=2e-7, atol=1e-6)
    env.close()


@pytest.mark.parametrize("math_mode", MATH_MODES)
@pytest.mark.parametrize("name", golden_names("bare_"))
def test_bare_ship_rollout_matches_reference_golden(name, math_mode):
    """KAT1 / KAT3: bare ship + controllers loop over 10k (simple) / 4k (detailed) steps."""
    g = golden(name)
    meta = json.loads(str(g["meta"]))
    if meta["kind"] == "simplified":
        # A8' (SURVEY.md section 8a): hull + SimplifiedMachineryModel, thrust-force state in the omega column
        args = S.get_env_args(time_step=meta["dt"])
        assets, m = S.build_simplified_assets(
            args, thrust_force_dynamic_time_constant=meta["thrust_force_dynamic_time_constant"],
            initial_thrust_force=meta["initial_thrust_force"], kp=meta["kp"], ki=meta["ki"])
        env = S.MultiShipRLEnv(assets=assets, map=m, args=args, math_mode=math_mode)
        assert env._params.ship[0].model_kind == L.MODEL_SIMPLIFIED
        done = 0
        for row, step in enumerate(g["step_index"]):
            env.ship_rollout(int(step) - done)
            done = int(step)
            if row % 8 == 0 or row == len(g["step_index"]) - 1:
                e = rel_err(product_ship_vec(env, 0), g["states"][row], STATE_SCALE)
                assert e.max() < REL_TOL, (name, row, e)
                assert int(env.next_wpt[0, 0]) == g["next_wpt"][done - 1]
        assert rel_err(float(env.ship_f64[L.SF["spd_err_i"], 0]), float(g["err_i"]), 1.0) < REL_TOL
        # the env level runs with this model too: a full episode terminates
        env.reset()
        for _ in range(9):
            o, r, d, info = env.step(np.array([0.01]))
            if d:
                break
        assert d and np.isfinite(r)
        env.close()
        return
    kind = "colav" if meta["kind"] == "simple" else "rl"
    env, _ = env_from_meta(dict(kind="noniw" if kind == "colav" else "rl", dt=meta["dt"], mode=meta.get("mode", "PTI")),
                           math_mode=math_mode)
    if meta["post_reset"] and kind == "rl":
        # reset() leaves the machinery at dt_shaft = 0.01; rebuild the un-stepped initial state after it
        env.reset()
        L.check(L.load().shipenv_construct(env._handle, None, env._stream_ptr()))
    role = meta["who"]
    steps = g["step_index"]
    detailed = kind == "rl"
    cfg = struct_from_bytes(O.ShipConfig, g["cfg"])
    tol = 1e-7 if meta["dt"] == 30 else REL_TOL
    if detailed:
        end_n, end_e = cfg.wp_north[cfg.n_wp - 1], cfg.wp_east[cfg.n_wp - 1]
        d_end = np.hypot(g["states"][:, 0] - end_n, g["states"][:, 1] - end_e)
        arrived = np.nonzero(d_end < 300.0)[0]
        last_row = arrived[0] if len(arrived) else len(d_end) - 1
    else:
        last_row = len(steps) - 1
    done = 0
    for row in range(last_row + 1):
        env.ship_rollout(int(steps[row]) - done)
        done = int(steps[row])
        if row % 8 == 0 or row == last_row:
            e = rel_err(product_ship_vec(env, role), g["states"][row], STATE_SCALE)
            assert e.max() < tol, (name, row, e)
            assert int(env.next_wpt[0, role]) == g["next_wpt"][done - 1]
    env.close()


# ------------------------------------------------------------------------------------------------
# batched runs against the CPU oracle on the same seeded inputs
# ------------------------------------------------------------------------------------------------
def _oracle_cfg(assets, env, kind):
    return O.env_config_from_assets(assets, env.map, env.args, kind)


@pytest.mark.parametrize("math_mode", MATH_MODES)
@pytest.mark.parametrize("fixture", REFERENCE_BATCHES)
def test_batched_rl_episodes_match_reference(fixture, math_mode):
    """Detailed model (config 3's env): 256 (collav none) / 64 (sbmpc, simple) jittered MultiShipRLEnv episodes,
    each compared with the UNMODIFIED REFERENCE's own run of the same seeded inputs (tests/golden/batch_rl_*.npz,
    made by tests/golden/make_reference_twins.py).

    Flags, event bits, step counts and waypoint indices are bit-exact; a flag may only differ where a one-ulp twin of
    the reference itself flips it (none does in these batches, so zero are waived).  States meet 1e-9, widened to
    min(1e-6, 10 x envelope) only where the reference's own one-ulp twins of that episode drift apart by the
    envelope recorded in the fixture (the cascaded throttle controller with measured_shaft_speed = forward_speed has
    gain ~1e4, rl_env controllers.py:185-189, env.py:397-401); the widened count is printed and bounded by the
    number of (environment, call) points whose reference envelope exceeds 1e-10."""
    g = golden(fixture)
    meta = json.loads(str(g["meta"]))
    B = meta["B"]
    args = S.get_env_args(time_step=4, collav_mode=meta["collav"])
    assets, _ = S.build_rl_assets(args)
    init = S.jittered_init_states(assets, B, pos_jitter_m=meta["pos_jitter_m"], seed=meta["seed_init"])
    assert np.array_equal(init.cpu().numpy().reshape(7, B, 2), g["init"])        # same seeded inputs as the fixture
    env, assets = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init, math_mode=math_mode)
    actions = torch.from_numpy(g["actions"])
    env.reset()
    alive = np.ones(B, dtype=bool)
    n_log = np.ones(B, dtype=np.int64)                      # rows of the reference's log: init_step + one per _step()
    stats = dict(waived_flags=[], waived_states=[], worst_tight=0.0)
    seen_events = 0
    for j in range(9):
        env.step(actions[:, j].cuda())
        _sync()
        info = env.info_buf.cpu().numpy()
        nsub = env.nsub_buf.cpu().numpy()
        obs = env.obs_buf.cpu().numpy()
        rew = env.reward_buf.cpu().numpy()
        states = product_states_all(env)
        kk = env.next_wpt.cpu().numpy()
        n_log += nsub
        for b in range(B):
            if not alive[b]:
                continue
            assert j < g["n_valid"][b]
            flags = (bool(info[b] & L.INFO_DONE), int(info[b] & L.INFO_EVENT_MASK), bool(info[b] & L.INFO_TERMINAL),
                     bool(info[b] & L.INFO_TEST_STOP), bool(info[b] & L.INFO_OBS_STOP), int(n_log[b]), int(kk[b, 0]),
                     int(kk[b, 1]))
            if not compare_with_reference_batch(g, b, j, states[b], flags, rew[b], obs[b], stats):
                alive[b] = False
                continue
            seen_events |= flags[1]
            if flags[0]:
                alive[b] = False
    assert not alive.any()
    n_ill = int((g["env_state"] * 10 > REL_TOL).sum())
    print(f"[{fixture}/{math_mode}] GPU vs reference: worst error of the calls held to 1e-9: {stats['worst_tight']:.2e}; "
          f"calls inside the reference's one-ulp envelope only: {len(stats['waived_states'])} of {n_ill} eligible "
          f"{[(b, j, f'{e:.1e}', f'{v:.1e}') for b, j, e, v in stats['waived_states']]}; waived flags: {stats['waived_flags']}")
    assert not stats["waived_flags"]
    assert len(stats["waived_states"]) <= n_ill
    assert bin(seen_events).count("1") >= 4, bin(seen_events)
    env.close()


@pytest.mark.parametrize("math_mode", MATH_MODES)
@pytest.mark.parametrize("collav", ["none", "simple", "sbmpc"])
@pytest.mark.parametrize("iw", [True, False], ids=["MultiShipEnv", "MultiShipNonIWEnv"])
def test_batched_colav_episodes_match_oracle(iw, collav, math_mode):
    """Simple model (config 2's env): 256 environments, per-env random scoping angles and jittered start positions,
    full episodes (9 step() calls); every environment is compared with its own scalar oracle run -- flags bit-exact,
    states within 1e-9, no allowance of any kind.  Both run_colav env classes: MultiShipEnv, and MultiShipNonIWEnv
    driven with step(action) (run_colav/env.py:678-800: collision avoidance on both ships, no travel tracker,
    six-entry observation)."""
    B = 256
    args = S.get_env_args(time_step=4, collav_mode=collav)
    route = dict() if iw else dict(obs_route="obs_ship_route.txt")
    assets, m = S.build_colav_assets(args, iw=iw, **route)
    init = S.jittered_init_states(assets, B, pos_jitter_m=100.0, seed=1)
    env, assets = S.prepare_colav_env(args, iw=iw, num_envs=B, init_states=init, math_mode=math_mode, **route)
    gen = torch.Generator().manual_seed(0)
    actions = (torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    actions[: B // 4] *= 0.1          # small angles keep the ships on a collision course
    env.reset()
    init_np = init.cpu().numpy().reshape(7, B, 2)
    base_cfg = _oracle_cfg(assets, env, O.ENV_COLAV_IW if iw else O.ENV_COLAV_NONIW)
    oracles = []
    for b in range(B):
        cfg = O.EnvConfig()
        C.memmove(C.byref(cfg), C.byref(base_cfg), C.sizeof(O.EnvConfig))
        for role in range(2):
            cfg.ship[role].initial_north_position_m = init_np[0, b, role]
            cfg.ship[role].initial_east_position_m = init_np[1, b, role]
        oe = O.OracleEnv(cfg)
        oe.reset()
        oracles.append(oe)
    alive = np.ones(B, dtype=bool)
    seen_events = 0
    worst = 0.0
    for j in range(9):
        env.step(actions[:, j].cuda())
        _sync()
        info = env.info_buf.cpu().numpy()
        nsub = env.nsub_buf.cpu().numpy()
        obs = env.obs_buf.cpu().numpy()
        states = product_states_all(env)
        kk = env.next_wpt.cpu().numpy()
        for b in range(B):
            if not alive[b]:
                continue
            r = oracles[b].step(float(actions[b, j]))
            st = oracles[b].st
            assert r.error == 0
            err = max(rel_err(states[b, role], oracle_ship_vec(st.ship[role]), STATE_SCALE).max() for role in range(2))
            assert (nsub[b] == r.n_substeps and (info[b] & L.INFO_EVENT_MASK) == r.events
                    and bool(info[b] & L.INFO_DONE) == bool(r.done)
                    and bool(info[b] & L.INFO_TERMINAL) == bool(r.terminal)
                    and bool(info[b] & L.INFO_TEST_STOP) == bool(r.test_ship_stop)
                    and bool(info[b] & L.INFO_OBS_STOP) == bool(r.obs_ship_stop)
                    and kk[b, 0] == st.ship[0].next_wpt and kk[b, 1] == st.ship[1].next_wpt), (b, j)
            assert err < REL_TOL, (b, j, err)
            worst = max(worst, err)
            np.testing.assert_allclose(obs[b], np.array(r.obs[:]), rtol=2e-7, atol=1e-6)
            seen_events |= r.events
            if r.done:
                alive[b] = False
                # finished environments are left alone by later calls
    assert not alive.any()
    print(f"[{'colav' if iw else 'noniw-step'}/{collav}/{math_mode}] worst rel err {worst:.2e}, events {seen_events:#x}")
    # the batch must have exercised several different endings (the NonIW class has no travel tracker, hence no
    # obstacle-ship navigation failure by distance travelled)
    assert bin(seen_events).count("1") >= (5 if iw else 3), bin(seen_events)
    env.close()


@pytest.mark.parametrize("kind", ["rl", "colav"])
def test_sbmpc_memory_survives_reset(kind):
    """SBMPCParams.P_ca_last_ / Chi_ca_last_ belong to the env's single SBMPC object: reset() does not
    clear them (sbmpc.py:30-31, env.py:123, 238-295), so the first SBMPC call of a second episode is
    penalised against the last manoeuvre of the first.  64 environments on a collision course, an
    abandoned episode followed by a full one, each environment compared with its own oracle."""
    B = 64
    n_first = 6 if kind == "rl" else 7      # step() calls of the first (abandoned) episode: the ships are passing each other
    args = S.get_env_args(time_step=4, collav_mode="sbmpc")
    if kind == "rl":
        assets, m = S.build_rl_assets(args)
        init = S.jittered_init_states(assets, B, pos_jitter_m=50.0, seed=11)
        env, assets = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init)
        okind = O.ENV_RL
    else:
        assets, m = S.build_colav_assets(args, iw=True)
        init = S.jittered_init_states(assets, B, pos_jitter_m=50.0, seed=11)
        env, assets = S.prepare_colav_env(args, iw=True, num_envs=B, init_states=init)
        okind = O.ENV_COLAV_IW
    gen = torch.Generator().manual_seed(4)
    actions = (torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6) * 0.08
    init_np = init.cpu().numpy().reshape(7, B, 2)
    base_cfg = _oracle_cfg(assets, env, okind)
    oracles = []
    for b in range(B):
        cfg = O.EnvConfig()
        C.memmove(C.byref(cfg), C.byref(base_cfg), C.sizeof(O.EnvConfig))
        for role in range(2):
            cfg.ship[role].initial_north_position_m = init_np[0, b, role]
            cfg.ship[role].initial_east_position_m = init_np[1, b, role]
        oracles.append(O.OracleEnv(cfg))
    carried = 0
    for episode in range(2):
        env.reset()
        for oe in oracles:
            oe.reset()
        if episode == 1:
            sb = env.env_f64[[L.EF["sb_p_last"], L.EF["sb_chi_last"]]].cpu().numpy()
            for b, oe in enumerate(oracles):
                assert sb[0, b] == oe.st.sb_p_last and sb[1, b] == oe.st.sb_chi_last
            carried = int(((sb[0] != 1.0) | (sb[1] != 0.0)).sum())
        alive = np.ones(B, dtype=bool)
        for j in range(n_first if episode == 0 else 9):
            env.step(actions[:, j].cuda())
            _sync()
            info = env.info_buf.cpu().numpy()
            nsub = env.nsub_buf.cpu().numpy()
            states = product_states_all(env)
            for b in range(B):
                if not alive[b]:
                    continue
                r = oracles[b].step(float(actions[b, j]))
                assert nsub[b] == r.n_substeps and (info[b] & L.INFO_EVENT_MASK) == r.events, (episode, b, j)
                if kind == "colav":     # the detailed model's conditioning is covered by the test above
                    for role in range(2):
                        e = rel_err(states[b, role], oracle_ship_vec(oracles[b].st.ship[role]), STATE_SCALE)
                        assert e.max() < REL_TOL, (episode, b, j, role, e)
                if r.done:
                    alive[b] = False
    assert carried > 0, "no environment ended its first episode in the middle of an SBMPC manoeuvre"
    env.close()


def test_substeps_split_invariance_and_masked_reset():
    """k x _step() in one launch == the same k steps in several launches (state stays in registers
    vs round-trips through HBM), bit for bit; masked reset only touches the selected environments."""
    B = 4096
    args = S.get_env_args(time_step=4)
    assets, _ = S.build_rl_assets(args)
    init = S.jittered_init_states(assets, B, seed=3)
    env_a, _ = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init)
    env_b, _ = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init)
    env_a.reset(); env_b.reset()
    env_a._step(96)
    for k in (1, 7, 24, 64):
        env_b._step(k)
    _sync()
    assert torch.equal(env_a.ship_f64, env_b.ship_f64)
    assert torch.equal(env_a.ship_i32, env_b.ship_i32)
    assert torch.equal(env_a.env_f64, env_b.env_f64)
    assert torch.equal(env_a.obs_buf, env_b.obs_buf)
    # (environments whose jittered start lies outside the map horizon finish early)
    assert env_a.total_substeps() == env_b.total_substeps() <= 96 * B
    assert env_a.total_substeps() > 48 * B
    # masked reset
    before = env_a.ship_f64.clone()
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    mask[::3] = True
    env_a.reset(mask=mask)
    _sync()
    s = env_a.ship_f64.view(L.SF_COUNT, B, 2)
    b4 = before.view(L.SF_COUNT, B, 2)
    assert torch.equal(s[:, ~mask], b4[:, ~mask])
    assert torch.all(s[L.SF["time"], mask] == 4.0)
    env_a.close(); env_b.close()


def test_full_size_properties_1e5_envs():
    """BASELINE config 2 size (1e5 SimpleShipModel pairs): identical environments stay identical, the
    device counters equal the per-env step counts, and every environment terminates."""
    B = 100_000
    args = S.get_env_args(time_step=4)
    env, assets = S.prepare_colav_env(args, iw=True, num_envs=B)
    env.reset()
    gen = torch.Generator().manual_seed(0)
    a64 = (torch.rand((64, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    actions = a64.repeat((B + 63) // 64, 1)[:B].cuda()      # env b uses action row b % 64
    total = 0
    for j in range(9):
        env.step(actions[:, j])
        total += int(env.nsub_buf.sum().item())
    _sync()
    assert env.total_substeps() == total
    assert bool(env.done_mask.all())
    st = env.ship_f64.view(L.SF_COUNT, B, 2)
    # environments 64 apart received identical inputs -> bit-identical trajectories
    assert torch.equal(st[:, :64], st[:, 64:128])
    ref = st[:, :64].repeat(1, (B + 63) // 64, 1)[:, :B]
    assert torch.equal(st, ref)
    # and the first 64 match the scalar oracle
    cfg = O.env_config_from_assets(assets, env.map, env.args, O.ENV_COLAV_IW)
    states = product_states_all(env)
    for b in range(0, 64, 7):
        oe = O.OracleEnv(cfg)
        oe.reset()
        for j in range(9):
            r = oe.step(float(a64[b, j]))
            if r.done:
                break
        for role in range(2):
            e = rel_err(states[b, role], oracle_ship_vec(oe.st.ship[role]), STATE_SCALE)
            assert e.max() < REL_TOL, (b, role, e)
        assert (int(env.info_buf[b]) & L.INFO_EVENT_MASK) == r.events
    env.close()


def test_full_size_bare_rollout_1e6_ships_10k_steps_sample():
    """BASELINE config 3 size: 1e6 detailed-model pairs integrate the bare loop; every ship of a role is
    bit-identical (same inputs), and the result matches the oracle after 1000 steps."""
    B = 1_000_000
    args = S.get_env_args(time_step=4)
    env, assets = S.prepare_multiship_rl_env(args, num_envs=B)
    env.ship_rollout(1000)
    _sync()
    st = env.ship_f64.view(L.SF_COUNT, B, 2)
    assert torch.equal(st[:, 1:], st[:, :1].expand(-1, B - 1, -1))
    for role in range(2):
        cfg = O.ship_config_from_asset(assets[role])
        out, wpt, _ = O.ship_rollout(cfg, 1000)
        e = rel_err(product_ship_vec(env, role, e=B - 1), out[-1], STATE_SCALE)
        assert e.max() < REL_TOL, (role, e)
        assert int(env.next_wpt[B - 1, role]) == wpt[-1]
    env.close()


def test_full_size_rl_episode_1e6_envs_jittered_with_reference_sample():
    """BASELINE config 3 at size: 1e6 jittered MultiShipRLEnv (ShipModelAST, PTI) environments run a whole episode with
    per-environment random scoping angles.  256 of them, scattered over the batch, are given the inputs of the
    reference-run fixture batch_rl_none_256 and are compared with the unmodified reference's results (flags bit-exact,
    states 1e-9 / reference envelope); the device step counter equals the sum of the per-call step counts; every
    environment terminates; neighbours of the sampled environments (different inputs) did not leak into them."""
    B = 1_000_000
    g = golden("batch_rl_none_256")
    nb = g["actions"].shape[0]
    idx = np.arange(nb) * (B // nb) + 17                      # where the fixture's environments sit in the batch
    args = S.get_env_args(time_step=4)
    assets, _ = S.build_rl_assets(args)
    init = S.jittered_init_states(assets, B, pos_jitter_m=50.0, seed=5, device="cpu").reshape(7, B, 2)
    init[:, idx, :] = torch.from_numpy(g["init"])
    gen = torch.Generator().manual_seed(9)
    actions = (torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    actions[idx] = torch.from_numpy(g["actions"])
    env, assets = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init.reshape(7, 2 * B).cuda())
    actions = actions.cuda()
    env.reset()
    idx_t = torch.from_numpy(idx).cuda()
    alive = np.ones(nb, dtype=bool)
    n_log = np.ones(nb, dtype=np.int64)
    stats = dict(waived_flags=[], waived_states=[], worst_tight=0.0)
    total = 0
    for j in range(9):
        env.step(actions[:, j])
        total += int(env.nsub_buf.sum(dtype=torch.int64).item())
        info = env.info_buf[idx_t].cpu().numpy()
        nsub = env.nsub_buf[idx_t].cpu().numpy()
        obs = env.obs_buf[idx_t].cpu().numpy()
        rew = env.reward_buf[idx_t].cpu().numpy()
        st = env.ship_f64.view(L.SF_COUNT, B, 2)[:, idx_t].cpu().numpy()
        states = np.stack([st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[8]], axis=-1)      # [nb, 2, 8]
        kk = env.next_wpt[idx_t].cpu().numpy()
        n_log += nsub
        for b in range(nb):
            if not alive[b]:
                continue
            flags = (bool(info[b] & L.INFO_DONE), int(info[b] & L.INFO_EVENT_MASK), bool(info[b] & L.INFO_TERMINAL),
                     bool(info[b] & L.INFO_TEST_STOP), bool(info[b] & L.INFO_OBS_STOP), int(n_log[b]), int(kk[b, 0]),
                     int(kk[b, 1]))
            ok = compare_with_reference_batch(g, b, j, states[b], flags, rew[b], obs[b], stats)
            if not ok or flags[0]:
                alive[b] = False
    _sync()
    assert not alive.any()
    assert not stats["waived_flags"]
    assert env.total_substeps() == total
    assert bool(env.done_mask.all())
    print(f"[1e6 rl] {total} simulator steps; sampled 256 vs reference: worst error held to 1e-9 {stats['worst_tight']:.2e}, "
          f"inside the reference envelope only: {len(stats['waived_states'])}")
    env.close()


def test_host_buffer_entry_points_match_device_path():
    """shipenv_reset_host / shipenv_step_host (numpy in, numpy out through the C ABI)."""
    B = 512
    args = S.get_env_args(time_step=4)
    env_d, _ = S.prepare_multiship_rl_env(args, num_envs=B)
    env_h, _ = S.prepare_multiship_rl_env(args, num_envs=B)
    gen = np.random.default_rng(5)
    actions = gen.uniform(-np.pi / 6, np.pi / 6, size=(B, 3))
    env_d.reset()
    obs0 = env_h.reset_host()
    assert np.array_equal(obs0, env_d.obs_buf.cpu().numpy())
    for j in range(3):
        env_d.step(torch.from_numpy(actions[:, j]).cuda())
        obs, rew, info, nsub = env_h.step_host(actions[:, j])
        _sync()
        assert np.array_equal(obs, env_d.obs_buf.cpu().numpy())
        assert np.array_equal(rew, env_d.reward_buf.cpu().numpy())
        assert np.array_equal(info, env_d.info_buf.cpu().numpy())
        assert np.array_equal(nsub, env_d.nsub_buf.cpu().numpy())
    env_d.close(); env_h.close()


def test_host_path_with_registered_arrays_matches_device_path():
    """The host-buffer path at a size where the env page-locks its arrays (shipenv_register_host: the copies go
    straight between them and the device).  Whole episode, then a masked reset and more calls, every output of every
    call against the device path."""
    B = 8192
    args = S.get_env_args(time_step=4)
    init = S.jittered_init_states(S.build_rl_assets(args)[0], B, pos_jitter_m=50.0, seed=21)
    env_d, _ = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init)
    env_h, _ = S.prepare_multiship_rl_env(args, num_envs=B, init_states=init)
    gen = np.random.default_rng(6)
    actions = gen.uniform(-np.pi / 6, np.pi / 6, size=(B, 12))
    def both(j):
        env_d.step(torch.from_numpy(actions[:, j]).cuda())
        obs, rew, info, nsub = env_h.step_host(actions[:, j])
        _sync()
        assert np.array_equal(obs, env_d.obs_buf.cpu().numpy()), j
        assert np.array_equal(rew, env_d.reward_buf.cpu().numpy()), j
        assert np.array_equal(info, env_d.info_buf.cpu().numpy()), j
        assert np.array_equal(nsub, env_d.nsub_buf.cpu().numpy()), j

    env_d.reset()
    assert np.array_equal(env_h.reset_host(), env_d.obs_buf.cpu().numpy())
    for j in range(9):
        both(j)
    assert bool(env_d.done_mask.all())
    mask = np.zeros(B, dtype=bool)
    mask[::3] = True
    env_d.reset(mask=torch.from_numpy(mask).cuda())
    assert np.array_equal(env_h.reset_host(mask), env_d.obs_buf.cpu().numpy())
    for j in range(9, 12):
        both(j)
    env_d.close(); env_h.close()


def test_c_abi_standalone_without_torch_buffers():
    """The library used as a plain C library: shipenv_alloc owns the device memory."""
    from ast_sac_b200 import env as E
    lib = L.load()
    args = S.get_env_args(time_step=4)
    assets, m = S.build_colav_assets(args, iw=True)
    P = E.pack_params(assets, m, args, L.ENV_COLAV_IW, post_reset=True)
    h = C.c_void_p()
    B = 1000
    L.check(lib.shipenv_create(C.byref(P), B, 0, C.byref(h)))
    assert lib.shipenv_step_host(h, None, None, None, None, None) != 0      # no buffers yet -> error code
    assert b"shipenv_bind" in lib.shipenv_last_error()
    L.check(lib.shipenv_alloc(h))
    obs = np.zeros((B, 8), np.float32)
    L.check(lib.shipenv_reset_host(h, None, obs.ctypes.data))
    assert np.allclose(obs[0], [100, 100, 0, 9900, 14900, -135 * np.pi / 180, 0, 3.5])
    actions = np.full(B, np.deg2rad(-2.0))
    rew = np.zeros(B); info = np.zeros(B, np.int32); nsub = np.zeros(B, np.int32)
    L.check(lib.shipenv_step_host(h, actions.ctypes.data, obs.ctypes.data, rew.ctypes.data, info.ctypes.data,
                                  nsub.ctypes.data))
    assert (nsub == 76).all()            # KAT2: 77 log rows after the first step() = init_step + 76 _step()
    cnt = (C.c_ulonglong * 4)()
    L.check(lib.shipenv_read_counters(h, cnt))
    assert cnt[0] == 76 * B
    L.check(lib.shipenv_destroy(h))


def test_step_after_budget_raises_like_reference():
    """Quirk 7: step() with the sampling budget exhausted and the obstacle ship
    inside a radius of acceptance leaves next_observations unbound in the reference."""
    args = S.get_env_args(time_step=4, max_sampling_frequency=1)
    env, _ = S.prepare_colav_env(args, iw=True)
    env.reset()
    o, d, info = env.step(np.array([0.0]))      # last sampling: runs to completion
    assert d
    env2, _ = S.prepare_colav_env(S.get_env_args(time_step=4, max_sampling_frequency=0), iw=True)
    env2.reset()
    # no sampling allowed: the loop runs until the obstacle ship reaches the RoA of the route end
    with pytest.raises(UnboundLocalError):
        env2.step(np.array([0.0]))
    env.close(); env2.close()


def test_noniw_step_needs_a_sampled_route_controller():
    """MultiShipNonIWEnv.step(action) calls auto_pilot.update_route (run_colav/env.py:602), which only
    HeadingBySampledRouteController has: with the plain HeadingByRouteController the reference raises AttributeError,
    and the C ABI refuses the call (SHIPENV_E_STATE) instead of sampling into a route that cannot take waypoints."""
    from ast_sac_b200.sim.controllers import HeadingByRouteController, HeadingControllerGains
    args = S.get_env_args(time_step=4)
    assets, m = S.build_colav_assets(args, iw=False)
    assets[1].auto_pilot = HeadingByRouteController(
        S.get_data_path("obs_ship_route_nonIW.txt"), heading_controller_gains=HeadingControllerGains(kp=.65, ki=0.001, kd=50),
        los_parameters=S._los(args), time_step=args.time_step, max_rudder_angle=np.deg2rad(30))
    from ast_sac_b200.env import MultiShipNonIWEnv
    env = MultiShipNonIWEnv(assets=assets, map=m, args=args)
    env.reset()
    with pytest.raises(AttributeError):
        env.step(np.array([0.0]))
    a = torch.zeros(1, dtype=torch.float64, device="cuda")
    assert L.load().shipenv_step(env._handle, a.data_ptr(), None) == 3      # SHIPENV_E_STATE
    o, d, info = env._step()            # the class's own stepping is unaffected
    assert not d and len(o) == 6
    env.close()


def test_pickle_roundtrip_rebuilds_device_state():
    import pickle
    args = S.get_env_args(time_step=4)
    env, _ = S.prepare_multiship_rl_env(args)
    env.reset()
    env2 = pickle.loads(pickle.dumps(env))
    o = env2.reset()
    assert np.array_equal(o, env.initial_states)
    r1 = env.step(np.array([0.01]))
    r2 = env2.step(np.array([0.01]))
    assert np.array_equal(r1[0], r2[0]) and r1[1] == r2[1]
    env.close(); env2.close()


# ------------------------------------------------------------------------------------------------
# trajectory log: the reference's ship_model.simulation_results
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["colav", "rl"])
def test_trajectory_log_matches_reference_simulation_results(kind):
    """KAT episode with the per-step log switched on: row counts equal the reference's log length after every
    step() call, and the logged columns (states before the integration, rudder, thrust / shaft speed, e_ct,
    heading error, repeated rows of a stopped ship) match the reference's simulation_results."""
    g = golden(f"log_{kind}_dt4_kat")
    ep = golden("colav_iw_dt4_kat" if kind == "colav" else "rl_dt4_kat")
    meta = json.loads(str(g["meta"]))
    env, assets = env_from_meta(meta, num_envs=3)
    env.enable_trajectory_log(n_envs=2, capacity=2048)
    env.reset()
    for j, a in enumerate(g["actions"]):
        res = env.step(np.full(3, a))
        _sync()
        assert int(env.log_count[1]) == int(ep["n_log"][j]), (j, int(env.log_count[1]), int(ep["n_log"][j]))
        if bool(res[-2][0]):
            break
    for role, name in ((0, "test"), (1, "obs")):
        sr = env.simulation_results(role, env=0)
        assert len(sr['time [s]']) == int(g[f"{name}_n"])
        rows = g[f"{name}_rows"]
        # the same keys in the same order as the reference's simulation_results (27 for ShipModelAST: the 12 state /
        # controller columns and the 15 machinery bookkeeping columns derived on the host, ship_model.py:911-937)
        ref_keys = [k.split("|", 1)[1] for k in g.files if k.startswith(name + "|")]
        assert list(sr.keys()) == ref_keys, (list(sr.keys()), ref_keys)
        for key in sr:
            gk = f"{name}|{key}"
            got = np.asarray(sr[key])[rows]
            want = g[gk]
            scale = {'yaw rate [deg/sec]': 1e-3 * 180 / np.pi, 'fuel rate me [kg/s]': 1e-3, 'fuel rate hsg [kg/s]': 1e-3,
                     'fuel rate [kg/s]': 1e-3, 'fuel consumption me [kg]': 1e-3, 'fuel consumption hsg [kg]': 1e-3,
                     'fuel consumption [kg]': 1e-3}.get(key, 1.0)
            e = rel_err(got, want, scale)
            assert e.max() < (1e-8 if kind == "rl" else REL_TOL), (kind, name, key, e.max(), int(e.argmax()))
        # environment 1 logged the same episode; the whole-episode columns of the episode fixture agree row by row
        assert np.array_equal(env.trajectory(role, env=1), env.trajectory(role, env=0))
        assert rel_err(np.asarray(sr['north position [m]']), ep[f"{name}_log_north"], 1.0).max() < 1e-8
        assert rel_err(np.asarray(sr['cross track error [m]']), ep[f"{name}_log_ect"], 1.0).max() < 1e-8
    # the host mirror exposes it like the reference object does
    assert assets[1].ship_model.simulation_results['time [s]'][:3] == [0.0, 4.0, 8.0]
    with pytest.raises(RuntimeError):
        env.trajectory(0, env=2)
    env.reset()
    _sync()
    assert int(env.log_count[0]) == 1 and int(env.log_count[1]) == 1       # reset() starts a new log; init_step logs one row
    env.close()


def test_thrust_state_model_env_episodes_match_oracle():
    """A8' at the env level: 64 jittered MultiShipRLEnv episodes whose hulls are driven by the SimplifiedMachineryModel
    (thrust-force state T) and ThrottleFromSpeedSetPointSimplifiedPropulsion, each against its own oracle run -- flags
    bit-exact, states 1e-9.  (The reference cannot run this model inside its env: SimplifiedMachineryModel has no
    recorded initial state to reset() to and its throttle controller does not take the env's measured_shaft_speed
    argument; its dynamics are pinned by the bare-loop golden bare_simplified_dt4_test.)"""
    B = 64
    tau, t0, kp, ki = 30.0, 0.0, 3.0, 0.02
    args = S.get_env_args(time_step=4)
    assets, m = S.build_simplified_assets(args, thrust_force_dynamic_time_constant=tau, initial_thrust_force=t0, kp=kp, ki=ki)
    init = S.jittered_init_states(assets, B, pos_jitter_m=100.0, seed=31)
    env = S.MultiShipRLEnv(assets=assets, map=m, args=args, num_envs=B, init_states=init)
    assert env._params.ship[0].model_kind == L.MODEL_SIMPLIFIED
    gen = torch.Generator().manual_seed(32)
    actions = (torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6)
    O.set_simplified(tau, t0)
    try:
        base_cfg = O.env_config_from_assets(assets, env.map, env.args, O.ENV_RL)
        init_np = init.cpu().numpy().reshape(7, B, 2)
        oracles = []
        for b in range(B):
            cfg = O.EnvConfig()
            C.memmove(C.byref(cfg), C.byref(base_cfg), C.sizeof(O.EnvConfig))
            for role in range(2):
                cfg.ship[role].initial_north_position_m = init_np[0, b, role]
                cfg.ship[role].initial_east_position_m = init_np[1, b, role]
            oe = O.OracleEnv(cfg)
            oe.reset()
            oracles.append(oe)
        env.reset()
        alive = np.ones(B, dtype=bool)
        worst, seen = 0.0, 0
        for j in range(9):
            env.step(actions[:, j].cuda())
            _sync()
            info = env.info_buf.cpu().numpy(); nsub = env.nsub_buf.cpu().numpy(); rew = env.reward_buf.cpu().numpy()
            st = env.ship_f64.cpu().numpy().reshape(L.SF_COUNT, B, 2)
            states = np.stack([st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[8]], axis=-1)
            kk = env.next_wpt.cpu().numpy()
            for b in range(B):
                if not alive[b]:
                    continue
                r = oracles[b].step(float(actions[b, j]))
                osh = oracles[b].st.ship
                assert nsub[b] == r.n_substeps and (info[b] & L.INFO_EVENT_MASK) == r.events, (b, j)
                assert bool(info[b] & L.INFO_DONE) == bool(r.done) and kk[b, 0] == osh[0].next_wpt and kk[b, 1] == osh[1].next_wpt
                err = max(rel_err(states[b, role], oracle_ship_vec(osh[role]), np.array([1, 1, 1, 1, 1, 1e-3, 1e3, 1])).max()
                          for role in range(2))
                assert err < REL_TOL, (b, j, err)
                assert rel_err(rew[b], r.reward, 1e-3) < 1e-7
                worst = max(worst, err)
                seen |= r.events
                if r.done:
                    alive[b] = False
        assert not alive.any() and seen != 0
        print(f"[thrust-state model, env level] worst rel err {worst:.2e}")
    finally:
        O.set_simplified(None)
    env.close()


# ------------------------------------------------------------------------------------------------
# NonIW env (config 1) batched: _step() in chunks against per-environment oracle runs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("collav", ["none", "simple", "sbmpc"])
def test_noniw_batched_substeps_match_oracle(collav):
    """64 MultiShipNonIWEnv copies with jittered starts, init_step() + _step(k) launches of uneven k until every
    environment is done; step counts, events and flags bit-exact against the oracle, states within REL_TOL.
    Under SBMPC both ships call the collision avoidance in turn (the obstacle ship's call sees the state the test
    ship has just integrated to): the two-phase path of the kernel."""
    B = 64
    args = S.get_env_args(time_step=4, collav_mode=collav)
    assets, m = S.build_colav_assets(args, iw=False)
    init = S.jittered_init_states(assets, B, pos_jitter_m=80.0, seed=5)
    env, assets = S.prepare_colav_env(args, iw=False, num_envs=B, init_states=init)
    base_cfg = _oracle_cfg(assets, env, O.ENV_COLAV_NONIW)
    init_np = init.cpu().numpy().reshape(7, B, 2)
    oracles, n_or, res = [], np.zeros(B, dtype=np.int64), [None] * B
    for b in range(B):
        cfg = O.EnvConfig()
        C.memmove(C.byref(cfg), C.byref(base_cfg), C.sizeof(O.EnvConfig))
        for role in range(2):
            cfg.ship[role].initial_north_position_m = init_np[0, b, role]
            cfg.ship[role].initial_east_position_m = init_np[1, b, role]
        oe = O.OracleEnv(cfg)
        oe.init_step()
        oracles.append(oe)
    env.init_step()
    total, chunk = 0, [1, 7, 64, 300, 129]
    it = 0
    while not bool(env.done_mask.all()) and total < 6000:
        k = chunk[it % len(chunk)]
        it += 1
        env._step(k)
        _sync()
        total += k
        nsub = env.nsub_buf.cpu().numpy()
        info = env.info_buf.cpu().numpy()
        states = product_states_all(env)
        kk = env.next_wpt.cpu().numpy()
        for b in range(B):
            for _ in range(int(nsub[b])):
                res[b] = oracles[b]._step()
                n_or[b] += 1
            if nsub[b] == 0:
                continue
            r, st = res[b], oracles[b].st
            assert (info[b] & L.INFO_EVENT_MASK) == r.events and bool(info[b] & L.INFO_DONE) == bool(r.done), (collav, b, total)
            assert bool(info[b] & L.INFO_TERMINAL) == bool(r.terminal)
            assert kk[b, 0] == st.ship[0].next_wpt and kk[b, 1] == st.ship[1].next_wpt
            for role in range(2):
                e = rel_err(states[b, role], oracle_ship_vec(st.ship[role]), STATE_SCALE)
                assert e.max() < REL_TOL, (collav, b, total, role, e)
            # an environment that is done stops early inside the launch: the oracle stopped at the same step
            assert bool(r.done) or nsub[b] == k
    assert bool(env.done_mask.all())
    assert env.total_substeps() == int(n_or.sum())
    env.close()


def test_substeps_split_invariance_with_sbmpc():
    """SBMPC keeps state between simulator steps (P_ca_last_, Chi_ca_last_): k x _step() in one launch must equal the
    same steps over several launches bit for bit, with the memory round-tripping through HBM."""
    B = 2048
    args = S.get_env_args(time_step=4, collav_mode="sbmpc")
    assets, _ = S.build_colav_assets(args, iw=True)
    # start the obstacle ship 100-250 m ahead of the test ship, head-on: SBMPC manoeuvres from the first step
    init = S.jittered_init_states(assets, B, pos_jitter_m=60.0, seed=9)
    init_v = init.view(7, B, 2)
    gap = torch.linspace(70.0, 180.0, B, dtype=torch.float64, device=init.device)
    init_v[0, :, 1] = init_v[0, :, 0] + gap
    init_v[1, :, 1] = init_v[1, :, 0] + gap
    env_a, _ = S.prepare_colav_env(args, iw=True, num_envs=B, init_states=init)
    env_b, _ = S.prepare_colav_env(args, iw=True, num_envs=B, init_states=init)
    env_a.init_step(); env_b.init_step()
    env_a._step(60)
    for k in (1, 9, 20, 30):
        env_b._step(k)
    _sync()
    assert torch.equal(env_a.ship_f64, env_b.ship_f64) and torch.equal(env_a.env_f64, env_b.env_f64)
    sb = env_a.env_f64[[L.EF["sb_p_last"], L.EF["sb_chi_last"]]]
    assert bool(((sb[0] != 1.0) | (sb[1] != 0.0)).any()), "SBMPC never chose a manoeuvre: the scenario does not exercise it"
    env_a.close(); env_b.close()


@pytest.mark.parametrize("math_mode", MATH_MODES)
def test_quiet_steps_are_bit_identical_to_full_evaluation(math_mode, monkeypatch):
    """The env kernel skips the event tests of a simulator step while every lane of the warp is provably far from its
    thresholds (quiet steps, shipenv_kernels.cuh).  SHIPENV_QUIET=0 makes the same kernels evaluate every test at every
    step: both must agree bit for bit -- observations, rewards, info words and step counts after every step() call and
    every state row at the end -- on 50 000 environments (more than the resident lane pairs: refills), start positions
    jittered by 400 m (ships that start outside the horizon, next to the islands, next to each other), per-env random
    scoping angles with a quarter of the environments kept on a collision course; then the same for _step(k)
    launches of several lengths."""
    B = 50_000
    args = S.get_env_args(time_step=4)
    assets, _ = S.build_colav_assets(args, iw=True)
    init = S.jittered_init_states(assets, B, pos_jitter_m=400.0, seed=5)
    monkeypatch.setenv("SHIPENV_QUIET", "0")
    ref, _ = S.prepare_colav_env(args, iw=True, num_envs=B, init_states=init, math_mode=math_mode)
    monkeypatch.delenv("SHIPENV_QUIET")
    env, _ = S.prepare_colav_env(args, iw=True, num_envs=B, init_states=init, math_mode=math_mode)
    gen = torch.Generator().manual_seed(11)
    actions = ((torch.rand((B, 9), generator=gen, dtype=torch.float64) * 2 - 1) * (np.pi / 6))
    actions[: B // 4] *= 0.1
    actions = actions.cuda()

    def same(what):
        for name in ("obs_buf", "reward_buf", "info_buf", "nsub_buf"):
            a, b = getattr(env, name), getattr(ref, name)
            assert torch.equal(a, b), (what, name, int((a != b).sum().item()))

    def same_state(what):
        for name in ("ship_f64", "ship_i32", "env_f64", "env_i32", "iw_f64"):
            a, b = getattr(env, name), getattr(ref, name)
            # (NaN-free by construction; equal_nan is not needed and would hide a divergence)
            assert torch.equal(a, b), (what, name, int((a != b).sum().item()))

    env.reset(); ref.reset()
    seen_words = set()
    for j in range(9):
        env.step(actions[:, j]); ref.step(actions[:, j])
        _sync()
        same(f"step() call {j}")
        seen_words.update(env.info_buf.to(torch.int64).bitwise_and(L.INFO_EVENT_MASK).unique().tolist())
    same_state("after the episode")
    assert env.total_substeps() == ref.total_substeps() > 50 * B
    assert bool(env.done_mask.all())
    bits = 0
    for w in seen_words:
        bits |= int(w)
    assert bin(bits).count("1") >= 6, hex(bits)   # collisions, groundings, navigation failures, route ends, horizon ...
    # _step(k): launches of one step, a few steps, and many steps (k_left is the launch's own counter there)
    env.reset(); ref.reset()
    for k in (1, 1, 3, 17, 128, 1, 400, 64):
        env._step(k); ref._step(k)
        _sync()
        same(f"_step({k})")
    same_state("after the _step() launches")
    assert env.total_substeps() == ref.total_substeps()
    env.close(); ref.close()
