"""ctypes binding of the CPU oracle (oracle/shipsim_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (ast_sac_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_WP, MAX_POLY, MAX_VERT = 32, 16, 128
MODEL_SIMPLE, MODEL_DETAILED, MODEL_SIMPLIFIED = 0, 1, 2
HSG = {"MOTOR": 0, "GEN": 1, "OFF": 2}
ENV_COLAV_NONIW, ENV_COLAV_IW, ENV_RL = 0, 1, 2
COLLAV = {"none": 0, None: 0, "simple": 1, "sbmpc": 2}

EVENT_STRINGS = [  # get_env_info.py:143-202 / reward_function.py:204-262, env.py:684
    'Ships collision!',
    '|Ship under test experiences grounding!|',
    '|Ship under test suffers navigational failure!|',
    '|Obstacle ship experiences grounding!|',
    '|Obstacle ship suffers navigational failure!|',
    '|Ship under test reaches its final destination!|',
    '|Ship under test goes outside the map horizon!|',
    '|Obstacle ship reaches its final destination!|',
    '|Obstacle ship goes outside the map horizon!|',
    '|Simulation reaches its time limit|',
    '|Learning agent samples false intermediate waypoints!|',
]


def events_to_string(bits: int) -> str:
    return ''.join(s for i, s in enumerate(EVENT_STRINGS) if bits & (1 << i))


_D = C.c_double
_SHIP_DOUBLES = [
    "dead_weight_tonnage", "coefficient_of_deadweight_to_displacement", "bunkers", "ballast",
    "length_of_ship", "width_of_ship",
    "added_mass_coefficient_in_surge", "added_mass_coefficient_in_sway", "added_mass_coefficient_in_yaw",
    "mass_over_linear_friction_coefficient_in_surge", "mass_over_linear_friction_coefficient_in_sway",
    "mass_over_linear_friction_coefficient_in_yaw",
    "nonlinear_friction_coefficient__in_surge", "nonlinear_friction_coefficient__in_sway",
    "nonlinear_friction_coefficient__in_yaw",
    "current_velocity_component_from_north", "current_velocity_component_from_east",
    "wind_speed", "wind_direction",
    "initial_north_position_m", "initial_east_position_m", "initial_yaw_angle_rad",
    "initial_forward_speed_m_per_s", "initial_sideways_speed_m_per_s", "initial_yaw_rate_rad_per_s",
    "integration_step", "simulation_time",
    "rudder_angle_to_sway_force_coefficient", "rudder_angle_to_yaw_force_coefficient",
    "hotel_load", "main_engine_capacity", "electrical_capacity", "rated_speed_main_engine_rpm",
    "linear_friction_main_engine", "linear_friction_hybrid_shaft_generator",
    "gear_ratio_between_main_engine_and_propeller", "gear_ratio_between_hybrid_shaft_generator_and_propeller",
    "propeller_inertia", "propeller_speed_to_torque_coefficient", "propeller_diameter",
    "propeller_speed_to_thrust_force_coefficient",
    "initial_propeller_shaft_speed_rad_per_s", "dt_shaft",
    "spd_kp", "spd_kd", "spd_ki", "max_thrust",
    "kp_ship_speed", "ki_ship_speed", "kp_shaft_speed", "ki_shaft_speed", "max_shaft_speed",
    "initial_shaft_speed_integral_error", "ctrl_time_step",
    "hdg_kp", "hdg_kd", "hdg_ki", "max_rudder_angle",
    "radius_of_acceptance", "lookahead_distance", "integral_gain", "integrator_windup_limit",
    "desired_forward_speed",
]


class ShipConfig(C.Structure):
    _fields_ = [(n, _D) for n in _SHIP_DOUBLES] + [
        ("wp_north", _D * MAX_WP), ("wp_east", _D * MAX_WP),
        ("n_wp", C.c_int32), ("model_kind", C.c_int32), ("shaft_generator_state", C.c_int32), ("pad_", C.c_int32)]


class SimplifiedMachinery(C.Structure):
    _fields_ = [("thrust_force_dynamic_time_constant", _D), ("initial_thrust_force", _D)]


class Map(C.Structure):
    _fields_ = [("n_poly", C.c_int32), ("poly_start", C.c_int32 * (MAX_POLY + 1)),
                ("vert_e", _D * MAX_VERT), ("vert_n", _D * MAX_VERT)]


class EnvConfig(C.Structure):
    _fields_ = [("ship", ShipConfig * 2), ("map", Map), ("env_kind", C.c_int32), ("collav", C.c_int32),
                ("max_sampling_frequency", C.c_int32), ("pad_", C.c_int32), ("radius_of_acceptance", _D)]


class ShipState(C.Structure):
    _fields_ = [(n, _D) for n in (
        "north", "east", "yaw", "u", "v", "r", "omega", "time", "e_ct", "e_ct_int", "hdg_err_i", "hdg_prev_err",
        "spd_err_i", "spd_prev_err", "shaft_err_i", "log_e_ct", "log_north", "log_east", "log_prev_north",
        "log_prev_east", "last_rudder", "last_thrust")] + [
        ("wp_north", _D * MAX_WP), ("wp_east", _D * MAX_WP),
        ("n_wp", C.c_int32), ("next_wpt", C.c_int32), ("prev_wpt", C.c_int32), ("stop_flag", C.c_int32),
        ("n_log", C.c_int32), ("pad_", C.c_int32)]


class EnvState(C.Structure):
    _fields_ = [("ship", ShipState * 2)] + [(n, _D) for n in (
        "travel_dist", "travel_time", "accumulated_rewards", "n_base", "e_base", "ab_segment_length",
        "ab_north_segment_length", "ab_east_segment_length", "omega")] + [
        ("states", C.c_float * 8), ("next_observations", C.c_float * 8), ("initial_states", C.c_float * 8),
        ("snapshot_events", C.c_int32), ("snapshot_terminal", C.c_int32), ("snapshot_test_stop", C.c_int32),
        ("snapshot_obs_stop", C.c_int32), ("sampling_count", C.c_int32), ("tracker_active", C.c_int32),
        ("n_substeps", C.c_int64), ("sb_p_last", _D), ("sb_chi_last", _D), ("sb_active", C.c_int32),
        ("pad2_", C.c_int32)]


class StepResult(C.Structure):
    _fields_ = [("obs", C.c_float * 8), ("reward", _D), ("last_step_reward", _D), ("done", C.c_int32),
                ("events", C.c_int32), ("terminal", C.c_int32), ("test_ship_stop", C.c_int32),
                ("obs_ship_stop", C.c_int32), ("n_substeps", C.c_int32), ("error", C.c_int32), ("pad_", C.c_int32)]


_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "shipsim_oracle.c")
    hdr = os.path.join(_HERE, "shipsim_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        assert L.orc_sizeof_ship_config() == C.sizeof(ShipConfig)
        assert L.orc_sizeof_env_config() == C.sizeof(EnvConfig)
        assert L.orc_sizeof_ship_state() == C.sizeof(ShipState)
        assert L.orc_sizeof_env_state() == C.sizeof(EnvState)
        assert L.orc_sizeof_step_result() == C.sizeof(StepResult)
        L.orc_map_distance.restype = _D
        L.orc_map_distance.argtypes = [C.POINTER(Map), _D, _D]
        L.orc_map_contains.argtypes = [C.POINTER(Map), _D, _D]
        L.orc_env_step.argtypes = [C.POINTER(EnvConfig), C.POINTER(EnvState), _D, C.POINTER(StepResult)]
        L.orc_ship_rollout.argtypes = [C.POINTER(ShipConfig), C.POINTER(ShipState), C.c_int64, C.c_int,
                                       C.c_void_p, C.c_void_p]
        L.orc_simplified_rollout.argtypes = [C.POINTER(ShipConfig), C.POINTER(SimplifiedMachinery), C.POINTER(ShipState),
                                             C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_bench_episodes.restype = C.c_int64
        L.orc_bench_episodes.argtypes = [C.POINTER(EnvConfig), C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_ship_config(route, model_kind=MODEL_SIMPLE, shaft_generator_state="MOTOR", **fields) -> ShipConfig:
    c = ShipConfig()
    for k, v in fields.items():
        if k not in _SHIP_DOUBLES:
            raise KeyError(k)
        setattr(c, k, float(v))
    route = np.asarray(route, dtype=np.float64).reshape(-1, 2)
    assert 2 <= len(route) <= MAX_WP
    c.n_wp = len(route)
    for i, (n, e) in enumerate(route):
        c.wp_north[i] = n
        c.wp_east[i] = e
    c.model_kind = model_kind
    c.shaft_generator_state = HSG[shaft_generator_state] if isinstance(shaft_generator_state, str) else shaft_generator_state
    if c.dt_shaft == 0.0:
        c.dt_shaft = c.integration_step
    if c.ctrl_time_step == 0.0:
        c.ctrl_time_step = c.integration_step
    return c


def make_map(map_data) -> Map:
    m = Map()
    m.n_poly = len(map_data)
    k = 0
    for p, poly in enumerate(map_data):
        m.poly_start[p] = k
        for (e, n) in poly:
            m.vert_e[k] = float(e)
            m.vert_n[k] = float(n)
            k += 1
    m.poly_start[len(map_data)] = k
    assert k <= MAX_VERT and len(map_data) <= MAX_POLY
    return m


def make_env_config(test: ShipConfig, obs: ShipConfig, map_data, env_kind, collav="none",
                    max_sampling_frequency=9, radius_of_acceptance=300.0) -> EnvConfig:
    cfg = EnvConfig()
    C.memmove(C.byref(cfg.ship[0]), C.byref(test), C.sizeof(ShipConfig))
    C.memmove(C.byref(cfg.ship[1]), C.byref(obs), C.sizeof(ShipConfig))
    mp = make_map(map_data)
    C.memmove(C.byref(cfg.map), C.byref(mp), C.sizeof(Map))
    cfg.env_kind = env_kind
    cfg.collav = COLLAV[collav]
    cfg.max_sampling_frequency = max_sampling_frequency
    cfg.radius_of_acceptance = float(radius_of_acceptance)
    return cfg


class OracleEnv:
    """One environment stepped by the C oracle, mirroring reset()/init_step()/_step()/step()."""

    def __init__(self, cfg: EnvConfig):
        self.cfg = EnvConfig()
        C.memmove(C.byref(self.cfg), C.byref(cfg), C.sizeof(EnvConfig))
        self.st = EnvState()
        lib().orc_env_construct(C.byref(self.cfg), C.byref(self.st))

    def reset(self):
        lib().orc_env_reset(C.byref(self.cfg), C.byref(self.st))
        return np.array(self.st.initial_states[:], dtype=np.float32)

    def init_step(self):
        lib().orc_env_init_step(C.byref(self.cfg), C.byref(self.st))

    def _step(self) -> StepResult:
        r = StepResult()
        lib().orc_env_substep(C.byref(self.cfg), C.byref(self.st), C.byref(r))
        return r

    def step(self, action: float) -> StepResult:
        r = StepResult()
        lib().orc_env_step(C.byref(self.cfg), C.byref(self.st), float(action), C.byref(r))
        return r

    def ship_state(self, who: int) -> np.ndarray:
        s = self.st.ship[who]
        return np.array([s.north, s.east, s.yaw, s.u, s.v, s.r, s.omega, s.e_ct], dtype=np.float64)


def ship_rollout(cfg: ShipConfig, n_steps: int, record_every: int = 1):
    """Bare ship + controllers loop; returns (states[n_rec, 8], next_wpt[n_rec], final ShipState)."""
    st = ShipState()
    lib().orc_ship_init(C.byref(cfg), C.byref(st))
    n_rec = n_steps // record_every
    out = np.zeros((n_rec, 8), dtype=np.float64)
    wpt = np.zeros((n_rec,), dtype=np.int32)
    lib().orc_ship_rollout(C.byref(cfg), C.byref(st), n_steps, record_every, out.ctypes.data, wpt.ctypes.data)
    return out, wpt, st


def simplified_rollout(cfg: ShipConfig, thrust_time_constant: float, initial_thrust: float, n_steps: int,
                       record_every: int = 1):
    """Bare loop of a hull driven by SimplifiedMachineryModel (cfg.model_kind == MODEL_SIMPLIFIED); the omega
    column of the returned states holds the thrust-force state."""
    assert cfg.model_kind == MODEL_SIMPLIFIED
    st = ShipState()
    lib().orc_ship_init(C.byref(cfg), C.byref(st))
    st.omega = initial_thrust                       # self.thrust = initial_thrust_force, ship_engine.py:504
    m = SimplifiedMachinery(thrust_time_constant, initial_thrust)
    n_rec = n_steps // record_every
    out = np.zeros((n_rec, 8), dtype=np.float64)
    wpt = np.zeros((n_rec,), dtype=np.int32)
    lib().orc_simplified_rollout(C.byref(cfg), C.byref(m), C.byref(st), n_steps, record_every, out.ctypes.data,
                                 wpt.ctypes.data)
    return out, wpt, st


def set_simplified(thrust_time_constant=None, initial_thrust: float = 0.0):
    """Machinery constants of the thrust-state model (A8') for OracleEnv runs whose ships are MODEL_SIMPLIFIED;
    None switches it off.  Global in the C library: set it before constructing the OracleEnv."""
    L = lib()
    L.orc_set_simplified.argtypes = [C.POINTER(SimplifiedMachinery)]
    L.orc_set_simplified.restype = None
    if thrust_time_constant is None:
        L.orc_set_simplified(None)
    else:
        m = SimplifiedMachinery(float(thrust_time_constant), float(initial_thrust))
        L.orc_set_simplified(C.byref(m))


def bench_episodes(cfg: EnvConfig, actions: np.ndarray, jitter_ne=None, n_threads: int = 0):
    """n_envs x n_rl_steps float64 actions -> (total _step() count, returns[n_envs], events[n_envs])."""
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    n_envs, n_rl = actions.shape
    ret = np.zeros(n_envs, dtype=np.float64)
    ev = np.zeros(n_envs, dtype=np.int32)
    jp = None
    if jitter_ne is not None:
        jitter_ne = np.ascontiguousarray(jitter_ne, dtype=np.float64)
        assert jitter_ne.shape == (n_envs, 2, 2)
        jp = jitter_ne.ctypes.data
    total = lib().orc_bench_episodes(C.byref(cfg), n_envs, n_rl, actions.ctypes.data, jp, n_threads,
                                     ret.ctypes.data, ev.ctypes.data)
    return int(total), ret, ev


# ------------------------------------------------------------------------------------------------
# config extraction from duck-typed asset objects (the reference's own objects in this container,
# or the product's host-side mirrors of them -- same attribute names)
# ------------------------------------------------------------------------------------------------
def ship_config_from_asset(asset, post_reset: bool = False) -> ShipConfig:
    sm = asset.ship_model
    f = {}
    f.update(sm.ship_config._asdict())
    f.update(sm.environment_config._asdict())
    f.update(sm.simulation_config._asdict())
    ap = asset.auto_pilot
    hc = ap.heading_controller
    pid = hc.ship_heading_controller
    f.update(hdg_kp=pid.kp, hdg_kd=pid.kd, hdg_ki=pid.ki, max_rudder_angle=hc.max_rudder_angle,
             ctrl_time_step=pid.time_step)
    nav = ap.navigate
    f.update(radius_of_acceptance=nav.ra, lookahead_distance=nav.r, integral_gain=nav.ki,
             integrator_windup_limit=nav.integrator_limit, desired_forward_speed=asset.desired_forward_speed)
    route = np.stack([np.asarray(nav.north, dtype=np.float64), np.asarray(nav.east, dtype=np.float64)], axis=1)
    if hasattr(sm, "ship_machinery_model") and hasattr(sm.ship_machinery_model, "thrust_time_constant"):
        # hull + SimplifiedMachineryModel (A8'): the machinery state slot holds the thrust force; the time constant
        # goes through set_simplified()
        mm = sm.ship_machinery_model
        tc = asset.throttle_controller
        f.update(rudder_angle_to_sway_force_coefficient=mm.c_rudder_v, rudder_angle_to_yaw_force_coefficient=mm.c_rudder_r,
                 hotel_load=mm.hotel_load, main_engine_capacity=mm.mode.main_engine_capacity,
                 electrical_capacity=mm.mode.electrical_capacity, initial_propeller_shaft_speed_rad_per_s=mm.thrust,
                 dt_shaft=mm.int.dt, kp_ship_speed=tc.ship_speed_controller.kp, ki_ship_speed=tc.ship_speed_controller.ki)
        return make_ship_config(route, model_kind=MODEL_SIMPLIFIED,
                                shaft_generator_state=mm.mode.shaft_generator_state, **f)
    if hasattr(sm, "ship_machinery_model"):
        mm = sm.ship_machinery_model
        tc = asset.throttle_controller
        f.update(rudder_angle_to_sway_force_coefficient=mm.c_rudder_v, rudder_angle_to_yaw_force_coefficient=mm.c_rudder_r,
                 hotel_load=mm.hotel_load, main_engine_capacity=mm.mode.main_engine_capacity,
                 electrical_capacity=mm.mode.electrical_capacity,
                 rated_speed_main_engine_rpm=mm.w_rated_me * 30 / np.pi,
                 linear_friction_main_engine=mm.d_me, linear_friction_hybrid_shaft_generator=mm.d_hsg,
                 gear_ratio_between_main_engine_and_propeller=mm.r_me,
                 gear_ratio_between_hybrid_shaft_generator_and_propeller=mm.r_hsg,
                 propeller_inertia=mm.jp, propeller_speed_to_torque_coefficient=mm.kp, propeller_diameter=mm.dp,
                 propeller_speed_to_thrust_force_coefficient=mm.kt,
                 initial_propeller_shaft_speed_rad_per_s=mm._initial_parameters['omega'] if hasattr(mm, '_initial_parameters') else mm.omega,
                 dt_shaft=0.01 if post_reset else mm.int.dt,
                 kp_ship_speed=tc.ship_speed_controller.kp, ki_ship_speed=tc.ship_speed_controller.ki,
                 kp_shaft_speed=tc.shaft_speed_controller.kp, ki_shaft_speed=tc.shaft_speed_controller.ki,
                 max_shaft_speed=tc.max_shaft_speed,
                 initial_shaft_speed_integral_error=tc.shaft_speed_controller._initial_state['error_i'])
        return make_ship_config(route, model_kind=MODEL_DETAILED,
                                shaft_generator_state=mm.mode.shaft_generator_state, **f)
    rc = sm.rudder_config
    sc = asset.speed_controller
    f.update(rudder_angle_to_sway_force_coefficient=rc.rudder_angle_to_sway_force_coefficient,
             rudder_angle_to_yaw_force_coefficient=rc.rudder_angle_to_yaw_force_coefficient,
             spd_kp=sc.ship_speed_controller.kp, spd_kd=sc.ship_speed_controller.kd,
             spd_ki=sc.ship_speed_controller.ki, max_thrust=sc.max_thrust)
    return make_ship_config(route, model_kind=MODEL_SIMPLE, **f)


def map_data_from_obstacle(map_obj):
    if hasattr(map_obj, "vertices"):      # product PolygonObstacle
        return map_obj.vertices
    out = []
    for p in map_obj.polygons:            # reference PolygonObstacle (stand-in or real shapely)
        coords = list(p.exterior.coords)
        if len(coords) > 1 and tuple(coords[0]) == tuple(coords[-1]):
            coords = coords[:-1]
        out.append([(float(x), float(y)) for x, y in coords])
    return out


def env_config_from_assets(assets, map_obj, args, env_kind) -> EnvConfig:
    test = ship_config_from_asset(assets[0])
    obs = ship_config_from_asset(assets[1])
    return make_env_config(test, obs, map_data_from_obstacle(map_obj), env_kind, collav=args.collav_mode,
                           max_sampling_frequency=args.max_sampling_frequency,
                           radius_of_acceptance=args.radius_of_acceptance)
