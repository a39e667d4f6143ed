"""Batched two-ship environments backed by the sm_100a kernels (csrc/) through the C ABI.

Drop-in mirrors of the reference's env classes -- same constructor ``(assets, map, args)``, same
``reset() / init_step() / _step() / step(action)`` methods, observation/action spaces and return
conventions -- with one new dimension: ``num_envs`` independent copies of the (test, obs) pair are
stepped by one kernel launch.

    reference class                                        here
    rl_env/ship_in_transit/env.py:41   MultiShipRLEnv      MultiShipRLEnv
    run_colav/env.py:810               MultiShipEnv        MultiShipEnv
    run_colav/env.py:37                MultiShipNonIWEnv   MultiShipNonIWEnv

With ``num_envs == 1`` (default) the methods return NumPy arrays / Python scalars / the env_info
dict exactly like the reference.  With ``num_envs > 1`` they return torch CUDA tensors (zero copy
views of the kernel's output buffers): ``obs [B, 8] float32``, ``reward [B] float64``,
``done [B] bool`` and an info dict of tensors (``events`` bit field, ``terminal``, ...).

There is no CPU path: constructing an env without the built extension or without a CUDA device raises.
"""
from __future__ import annotations

import copy
import ctypes as C
from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import List, Optional, Union

import numpy as np
import torch

from . import _lib as L
from .sim.controllers import (EngineThrottleFromSpeedSetPoint, HeadingByRouteController,
                              HeadingBySampledRouteController, ThrustFromSpeedSetPoint)
from .sim.obstacle import PolygonObstacle
from .sim.ship_model import BaseShipModel
from .spaces import Box

# "strict": the reference's formulas statement by statement, no FMA contraction.
# "fast":   algebraically identical rewrites with fewer transcendental calls + FMA contraction.
# Both pass the same parity tests (tests/test_gpu_parity.py); see DESIGN.md section 3.
DEFAULT_MATH_MODE = "fast"

EVENT_STRINGS = [  # reward_function.py:204-262 / get_env_info.py:143-202, env.py:684
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


class BatchedInfo(Mapping):
    """``env_info`` of a batched call: the reference's keys ('events', 'terminal', 'test_ship_stop',
    'obs_ship_stop'; reward_function.py:204-270) plus 'substeps', one entry per environment.

    The entries are decoded from the call's packed info words when they are first read (and kept): a step() that
    nobody asks for its env_info launches no decoding kernels -- at 1e5 environments the seven small kernels of
    an eagerly built dict were 4 % of a step() call.  Like ``obs`` and ``reward`` the words are the environment's
    output buffer: read the entries before the next step() / _step() / reset() overwrites it."""
    _KEYS = ('events', 'terminal', 'test_ship_stop', 'obs_ship_stop', 'substeps')

    def __init__(self, info_buf, nsub_buf):
        self._info, self._nsub, self._cache = info_buf, nsub_buf, {}

    def __getitem__(self, key):
        if key not in self._cache:
            info = self._info
            if key == 'events':
                v = info & L.INFO_EVENT_MASK
            elif key == 'terminal':
                v = (info & L.INFO_TERMINAL) != 0
            elif key == 'test_ship_stop':
                v = (info & L.INFO_TEST_STOP) != 0
            elif key == 'obs_ship_stop':
                v = (info & L.INFO_OBS_STOP) != 0
            elif key == 'substeps':
                v = self._nsub
            else:
                raise KeyError(key)
            self._cache[key] = v
        return self._cache[key]

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


def events_to_string(bits: int) -> str:
    """Event bit field -> the reference's concatenated ``env_info['events']`` string."""
    return ''.join(s for i, s in enumerate(EVENT_STRINGS) if bits & (1 << i))


@dataclass
class ShipAssets:
    """rl_env/ship_in_transit/env.py:29-39 and run_colav/env.py:25-35 (``speed_controller`` there)."""
    ship_model: BaseShipModel
    auto_pilot: Union[HeadingBySampledRouteController, HeadingByRouteController]
    desired_forward_speed: float
    integrator_term: List[float]
    time_list: List[float]
    type_tag: str
    stop_flag: bool
    throttle_controller: Optional[EngineThrottleFromSpeedSetPoint] = None
    speed_controller: Optional[ThrustFromSpeedSetPoint] = None
    init_copy: 'ShipAssets' = field(default=None, repr=False, compare=False)


# ------------------------------------------------------------------------------------------------
# parameter packing (host objects -> ShipEnvParams)
# ------------------------------------------------------------------------------------------------
def pack_ship_params(asset: ShipAssets, nav_fail_tol: float, dt_shaft: Optional[float] = None) -> L.ShipParams:
    sm = asset.ship_model
    p = L.ShipParams()
    p.mass, p.i_z, p.x_du, p.y_dv, p.n_dr = sm.mass, sm.i_z, sm.x_du, sm.y_dv, sm.n_dr
    p.lin_damp_u = sm.mass / sm.t_surge            # linear_damping_matrix, ship_model.py:212-215
    p.lin_damp_v = sm.mass / sm.t_sway
    p.lin_damp_r = sm.i_z / sm.t_yaw
    p.ku, p.kv, p.kr = sm.ku, sm.kv, sm.kr
    p.inv_m_u = 1.0 / (sm.mass + sm.x_du)          # inverse of the diagonal mass matrix (x_g = 0)
    p.inv_m_v = 1.0 / (sm.mass + sm.y_dv)
    p.inv_m_r = 1.0 / (sm.i_z + sm.n_dr)
    p.cur_n, p.cur_e = float(sm.vel_c[0]), float(sm.vel_c[1])
    p.wind_speed, p.wind_dir = sm.wind_speed, sm.wind_dir
    p.cos_wind_dir, p.sin_wind_dir = float(np.cos(sm.wind_dir)), float(np.sin(sm.wind_dir))
    p.proj_area_f, p.proj_area_l, p.l_ship = sm.proj_area_f, sm.proj_area_l, sm.l_ship
    p.w_ship = sm.w_ship
    sc = sm.simulation_config
    p.init_north, p.init_east, p.init_yaw = sc.initial_north_position_m, sc.initial_east_position_m, sc.initial_yaw_angle_rad
    p.init_u, p.init_v, p.init_r = (sc.initial_forward_speed_m_per_s, sc.initial_sideways_speed_m_per_s,
                                    sc.initial_yaw_rate_rad_per_s)
    p.dt, p.sim_time = sc.integration_step, sc.simulation_time
    ap = asset.auto_pilot
    pid = ap.heading_controller.ship_heading_controller
    p.ctrl_dt = pid.time_step
    p.inv_ctrl_dt = 1.0 / pid.time_step
    p.hdg_kp, p.hdg_kd, p.hdg_ki = pid.kp, pid.kd, pid.ki
    p.max_rudder = ap.heading_controller.max_rudder_angle
    nav = ap.navigate
    p.los_ra, p.los_r, p.los_ki, p.los_limit = nav.ra, nav.r, nav.ki, nav.integrator_limit
    p.desired_speed = asset.desired_forward_speed
    p.nav_fail_tol = nav_fail_tol
    n_wp = len(nav.north)
    if not (2 <= n_wp <= L.MAX_WP):
        raise ValueError(f"route must have between 2 and {L.MAX_WP} waypoints, got {n_wp}")
    for i in range(n_wp):
        p.wp_north[i] = float(nav.north[i])
        p.wp_east[i] = float(nav.east[i])
    p.n_wp = n_wp
    if hasattr(sm, "ship_machinery_model") and hasattr(sm.ship_machinery_model, "thrust_time_constant"):
        # hull + SimplifiedMachineryModel (thrust-force state), ship_engine.py:484-519
        mm = sm.ship_machinery_model
        tc = asset.throttle_controller
        if tc is None or not hasattr(tc, "ship_speed_controller") or hasattr(tc, "shaft_speed_controller"):
            raise ValueError("ShipModelSimplifiedPropulsion assets need a ThrottleFromSpeedSetPointSimplifiedPropulsion")
        p.model_kind = L.MODEL_SIMPLIFIED
        p.c_rudder_v, p.c_rudder_r = mm.c_rudder_v, mm.c_rudder_r
        p.init_omega = mm.thrust                       # the machinery state row holds the thrust force
        p.dt_shaft = mm.int.dt        # (the reference defines no reset() for this machinery model: no dt quirk)
        p.kp_ship_speed, p.ki_ship_speed = tc.ship_speed_controller.kp, tc.ship_speed_controller.ki
        if tc.ship_speed_controller.time_step != pid.time_step:
            raise ValueError("all controllers of an asset must share one time_step")
        p.p_me = mm.mode.available_propulsion_power_main_engine
        p.p_el = mm.mode.available_propulsion_power_electrical
        p.k_thrust, p.thrust_tau = mm.k_thrust, mm.thrust_time_constant
    elif hasattr(sm, "ship_machinery_model"):
        mm = sm.ship_machinery_model
        tc = asset.throttle_controller
        if tc is None:
            raise ValueError("ShipModelAST assets need a throttle_controller (EngineThrottleFromSpeedSetPoint)")
        p.model_kind = L.MODEL_DETAILED
        p.c_rudder_v, p.c_rudder_r = mm.c_rudder_v, mm.c_rudder_r
        p.init_omega = mm.omega
        p.dt_shaft = mm.int.dt if dt_shaft is None else dt_shaft
        p.kp_ship_speed, p.ki_ship_speed = tc.ship_speed_controller.kp, tc.ship_speed_controller.ki
        p.kp_shaft_speed, p.ki_shaft_speed = tc.shaft_speed_controller.kp, tc.shaft_speed_controller.ki
        p.max_shaft_speed = tc.max_shaft_speed
        p.init_shaft_err_i = tc.shaft_speed_controller._initial_state['error_i']
        if tc.ship_speed_controller.time_step != pid.time_step:
            raise ValueError("all controllers of an asset must share one time_step")
        p.p_me = mm.mode.available_propulsion_power_main_engine
        p.p_el = mm.mode.available_propulsion_power_electrical
        p.tq_me_max = mm.mode.available_propulsion_power_main_engine / 5 * np.pi / 30   # ship_engine.py:422-423
        p.tq_el_max = mm.mode.available_propulsion_power_electrical / 5 * np.pi / 30    # ship_engine.py:431-432
        p.d_me, p.d_hsg, p.r_me, p.r_hsg = mm.d_me, mm.d_hsg, mm.r_me, mm.r_hsg
        p.jp, p.k_torque = mm.jp, mm.kp
        p.thrust_coeff = mm.dp ** 4 * mm.kt                                             # ship_engine.py:414
    else:
        spd = asset.speed_controller
        if spd is None:
            raise ValueError("SimpleShipModel assets need a speed_controller (ThrustFromSpeedSetPoint)")
        p.model_kind = L.MODEL_SIMPLE
        rc = sm.rudder_config
        p.c_rudder_v, p.c_rudder_r = rc.rudder_angle_to_sway_force_coefficient, rc.rudder_angle_to_yaw_force_coefficient
        p.dt_shaft = sc.integration_step
        c = spd.ship_speed_controller
        p.spd_kp, p.spd_kd, p.spd_ki, p.max_thrust = c.kp, c.kd, c.ki, spd.max_thrust
        if c.time_step != pid.time_step:
            raise ValueError("all controllers of an asset must share one time_step")
    return p


def pack_params(assets, map_obj: PolygonObstacle, args, env_kind: int, post_reset: bool,
                math_mode: int = L.MATH_STRICT) -> L.Params:
    if len(assets) != 2:
        raise ValueError("assets must be [ship under test, obstacle ship]")
    P = L.Params()
    dt_shaft = 0.01 if post_reset else None        # BaseMachineryModel.reset quirk, ship_engine.py:331-333
    ps = [pack_ship_params(assets[0], 3000.0, dt_shaft), pack_ship_params(assets[1], 500.0, dt_shaft)]
    for i in range(2):
        C.memmove(C.byref(P.ship[i]), C.byref(ps[i]), C.sizeof(L.ShipParams))
    k = 0
    if map_obj.num_obstacles > L.MAX_POLY:
        raise ValueError(f"at most {L.MAX_POLY} polygons")
    for pi, poly in enumerate(map_obj.vertices):
        P.poly_start[pi] = k
        for (e, n) in poly:
            if k >= L.MAX_VERT:
                raise ValueError(f"at most {L.MAX_VERT} map vertices")
            P.vert_e[k] = e
            P.vert_n[k] = n
            k += 1
    P.poly_start[map_obj.num_obstacles] = k
    P.n_poly = map_obj.num_obstacles
    P.map_min_n, P.map_max_n = map_obj.min_north, map_obj.max_north
    P.map_min_e, P.map_max_e = map_obj.min_east, map_obj.max_east
    # init_get_intermediate_waypoints (rl_env env.py:143-162), evaluated with numpy like the reference
    nav = assets[1].auto_pilot.navigate
    msf = args.max_sampling_frequency
    ab_n = nav.north[-1] - nav.north[0]
    ab_e = nav.east[-1] - nav.east[0]
    ab_length = np.sqrt(ab_n ** 2 + ab_e ** 2)
    P.ab_segment_length = ab_length / (msf + 1)
    P.ab_north_segment_length = ab_n / (msf + 1)
    P.ab_east_segment_length = ab_e / (msf + 1)
    ab_alpha = np.arctan2(ab_e, ab_n)
    ab_beta = np.pi / 2 - ab_alpha
    omega = np.pi / 2 - ab_beta
    P.cos_omega, P.sin_omega = float(np.cos(omega)), float(np.sin(omega))
    P.n_base0 = P.ab_north_segment_length + nav.north[0]
    P.e_base0 = P.ab_east_segment_length + nav.east[0]
    P.roa = args.radius_of_acceptance
    P.env_kind = env_kind
    collav = args.collav_mode
    if collav in (None, 'none', 'None'):
        P.collav = L.COLLAV_NONE
    elif collav == 'simple':
        P.collav = L.COLLAV_SIMPLE
    elif collav == 'sbmpc':
        P.collav = L.COLLAV_SBMPC
    else:
        raise ValueError(f"unknown collav_mode {collav!r}")
    P.max_sampling_frequency = msf
    # MultiShipNonIWEnv.step(action) needs an obstacle ship whose autopilot can take intermediate waypoints
    # (auto_pilot.update_route, run_colav/env.py:602: only HeadingBySampledRouteController has it)
    P.obs_sampled_route = int(env_kind == L.ENV_COLAV_NONIW and isinstance(assets[1].auto_pilot, HeadingBySampledRouteController))
    P.abi_version = L.ABI_VERSION
    P.math_mode = math_mode
    return P


# ------------------------------------------------------------------------------------------------
# the batched env
# ------------------------------------------------------------------------------------------------
class BatchedShipEnv:
    ENV_KIND = L.ENV_RL
    OBS_DIM = 8

    def __init__(self, assets: List[ShipAssets], map: PolygonObstacle, args, num_envs: int = 1,
                 device: Union[None, int, str, torch.device] = None, init_states: Optional[torch.Tensor] = None,
                 math_mode: Union[str, int, None] = None):
        self.args = args
        if math_mode is None:
            math_mode = getattr(args, "math_mode", DEFAULT_MATH_MODE)
        self.math_mode = {"strict": L.MATH_STRICT, "fast": L.MATH_FAST}.get(math_mode, math_mode)
        if self.math_mode not in (L.MATH_STRICT, L.MATH_FAST):
            raise ValueError("math_mode must be 'strict' or 'fast'")
        self.collav = args.collav_mode
        self.assets = assets
        [self.test, self.obs] = self.assets
        for asset in self.assets:
            asset.init_copy = None
            asset.init_copy = copy.deepcopy(asset)
        self.map = map
        self.num_envs = int(num_envs)
        self.batched = self.num_envs > 1
        self.ship_draw = getattr(args, "ship_draw", False)
        self.time_since_last_ship_drawing = getattr(args, "time_since_last_ship_drawing", 30)

        # observation / action spaces: rl_env env.py:86-104, run_colav env.py:855-866
        self.observation_space = Box(
            low=np.array([0, 0, -3000, 0, 0, -np.pi, -3000, 0], dtype=np.float32),
            high=np.array([10000, 20000, 3000, 10000, 20000, np.pi, 3000, 10], dtype=np.float32), dtype=np.float32)
        self.obsv_dim = self.observation_space.shape[0]
        self.action_space = self._make_action_space(args)
        self.action_dim = self.action_space.shape[0]

        tm, om = self.test.ship_model, self.obs.ship_model
        self.initial_states = np.array([tm.north, tm.east, 0.0, om.north, om.east, om.yaw_angle, 0.0,
                                        om.forward_speed], dtype=np.float32)

        self._post_reset = False
        self._handle = None
        self._device = self._resolve_device(device)
        self._init_states = None
        if init_states is not None:
            t = torch.as_tensor(init_states, dtype=torch.float64, device=self._device).contiguous()
            if tuple(t.shape) != (7, 2 * self.num_envs):
                raise ValueError("init_states must have shape [7, 2 * num_envs] (north, east, yaw, u, v, r, omega)")
            self._init_states = t
        self._open()
        for role, asset in enumerate(self.assets):
            asset.ship_model._binding = (self, role)

    # -- construction helpers ---------------------------------------------------------------------
    def _make_action_space(self, args):
        return Box(low=np.array([-np.pi / 6], dtype=np.float32), high=np.array([np.pi / 6], dtype=np.float32),
                   dtype=np.float32)

    @staticmethod
    def _resolve_device(device) -> torch.device:
        if not torch.cuda.is_available():
            raise RuntimeError("ast_sac_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device is None:
            return torch.device("cuda", torch.cuda.current_device())
        d = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if d.type != "cuda":
            raise ValueError("device must be a CUDA device")
        return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())

    def _open(self):
        lib = L.load()
        self._params = pack_params(self.assets, self.map, self.args, self.ENV_KIND, self._post_reset, self.math_mode)
        h = C.c_void_p()
        L.check(lib.shipenv_create(C.byref(self._params), self.num_envs, self._device.index, C.byref(h)))
        self._handle = h
        lay = L.Layout()
        L.check(lib.shipenv_layout(h, C.byref(lay)))
        B, dev = self.num_envs, self._device
        f64, i32, f32 = torch.float64, torch.int32, torch.float32
        self.ship_f64 = torch.zeros((L.SF_COUNT, 2 * B), dtype=f64, device=dev)
        self.ship_i32 = torch.zeros((2 * B,), dtype=i32, device=dev)
        self.env_f64 = torch.zeros((L.EF_COUNT, B), dtype=f64, device=dev)
        self.env_i32 = torch.zeros((L.EI_COUNT, B), dtype=i32, device=dev)
        self.iw_f64 = torch.zeros((2, L.MAX_IW, B), dtype=f64, device=dev)
        self.prev_f32 = torch.zeros((4, B), dtype=f32, device=dev)
        # the four per-call outputs back to back (obs | reward | info | nsub, 48 B per environment): the host-buffer
        # path then fetches them with one copy
        self._out_all = torch.zeros((48 * B,), dtype=torch.uint8, device=dev)
        self.obs_buf = self._out_all[:32 * B].view(f32).view(B, 8)
        self.reward_buf = self._out_all[32 * B:40 * B].view(f64)
        self.info_buf = self._out_all[40 * B:44 * B].view(i32)
        self.nsub_buf = self._out_all[44 * B:48 * B].view(i32)
        self.counters = torch.zeros((4,), dtype=torch.int64, device=dev)
        assert self.ship_f64.numel() == lay.ship_f64 and self.iw_f64.numel() == lay.iw_f64
        assert self.env_f64.numel() == lay.env_f64 and self.env_i32.numel() == lay.env_i32
        b = L.Buffers(self.ship_f64.data_ptr(), self.ship_i32.data_ptr(), self.env_f64.data_ptr(),
                      self.env_i32.data_ptr(), self.iw_f64.data_ptr(), self.prev_f32.data_ptr(),
                      self.obs_buf.data_ptr(), self.reward_buf.data_ptr(), self.info_buf.data_ptr(),
                      self.nsub_buf.data_ptr(), self.counters.data_ptr())
        L.check(lib.shipenv_bind(h, C.byref(b)))
        init_ptr = self._init_states.data_ptr() if self._init_states is not None else None
        L.check(lib.shipenv_construct(h, init_ptr, self._stream_ptr()))
        if self._init_states is not None:
            self.initial_states_batch = self.obs_buf.clone()

    def _stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def close(self):
        if getattr(self, "_handle", None) is not None:
            torch.cuda.synchronize(self._device)
            L.load().shipenv_destroy(self._handle)      # (releases the host registrations of _host_arrays)
            self._handle = None
            self.__dict__.pop("_host_bufs", None)
            self.__dict__.pop("_host_actions", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # the reference pickles the whole env into training snapshots (path_collector.py:96-102)
    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if k not in (
            "_handle", "_params", "ship_f64", "ship_i32", "env_f64", "env_i32", "iw_f64", "prev_f32", "obs_buf",
            "reward_buf", "info_buf", "nsub_buf", "_out_all", "counters", "log_f64", "log_count", "_log_envs", "_host_bufs",
            "_host_actions")}
        d["_device"] = str(self._device)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._device = self._resolve_device(self._device)
        self._handle = None
        self._open()
        for role, asset in enumerate(self.assets):
            asset.ship_model._binding = (self, role)

    # -- helpers ----------------------------------------------------------------------------------
    def _actions_tensor(self, action) -> torch.Tensor:
        if isinstance(action, torch.Tensor):
            a = action.to(device=self._device, dtype=torch.float64)
        else:
            a = torch.as_tensor(np.asarray(action, dtype=np.float64), device=self._device)
        a = a.reshape(-1)
        if a.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {a.numel()}")
        return a.contiguous()

    def _info_tensors(self):
        return BatchedInfo(self.info_buf, self.nsub_buf)

    def _info_dict_scalar(self, bits: int):
        return {
            'events': events_to_string(bits & L.INFO_EVENT_MASK),
            'terminal': bool(bits & L.INFO_TERMINAL),
            'test_ship_stop': bool(bits & L.INFO_TEST_STOP),
            'obs_ship_stop': bool(bits & L.INFO_OBS_STOP),
        }

    def _obs_numpy(self) -> np.ndarray:
        return self.obs_buf[0, :self.OBS_DIM].cpu().numpy()

    def read_ship_state(self, role: int, env: int = 0) -> np.ndarray:
        """[north, east, yaw, u, v, r, omega, time, e_ct, e_ct_int, ...] of one ship (device read)."""
        return self.ship_f64[:, 2 * env + role].cpu().numpy()

    @property
    def ship_state(self) -> torch.Tensor:
        """Per-ship FP64 state as a view [SF_COUNT, num_envs, 2] (last axis: test, obs)."""
        return self.ship_f64.view(L.SF_COUNT, self.num_envs, 2)

    @property
    def sampling_count(self):
        sc = self.env_i32[L.EI["sampling_count"]]
        return sc if self.batched else int(sc[0])

    @property
    def next_wpt(self) -> torch.Tensor:
        return (self.ship_i32 & 0xff).view(self.num_envs, 2)

    @property
    def stop_flags(self) -> torch.Tensor:
        return ((self.ship_i32 >> 8) & 1).view(self.num_envs, 2)

    @property
    def done_mask(self) -> torch.Tensor:
        return (self.env_i32[L.EI["flags"]] & 1) != 0

    def total_substeps(self) -> int:
        """Number of _step() calls executed since construction, summed over environments."""
        return int(self.counters[0].item())

    def obs_route(self, env: int = 0):
        """Obstacle ship route of one environment, including the sampled intermediate waypoints."""
        nav = self.obs.auto_pilot.navigate
        sampled = self.ENV_KIND != L.ENV_COLAV_NONIW or self._params.obs_sampled_route
        n_iw = int(self.env_i32[L.EI["sampling_count"], env]) if sampled else 0
        iw = self.iw_f64[:, :n_iw, env].cpu().numpy()
        north = [float(x) for x in nav.north[:-1]] + iw[0].tolist() + [float(nav.north[-1])]
        east = [float(x) for x in nav.east[:-1]] + iw[1].tolist() + [float(nav.east[-1])]
        return north, east

    def set_done(self, mask):
        """Mark the masked environments as finished: step() / _step() leave them untouched (and they cost the
        kernel one queue fetch) until they are reset.  Used by the vectorised sampler to park the environments
        it does not need for the current wave."""
        m = torch.as_tensor(mask, device=self._device).to(torch.bool).reshape(-1)
        if m.numel() != self.num_envs:
            raise ValueError("mask must have num_envs entries")
        flags = self.env_i32[L.EI["flags"]]
        flags.copy_(torch.where(m, flags | 1, flags))

    # -- trajectory log ----------------------------------------------------
    def enable_trajectory_log(self, n_envs: int = 1, capacity: int = 4096):
        """Record the per-step rows the reference keeps in ``ship_model.simulation_results``
        (store_simulation_data / store_last_simulation_data, ship_model.py:418-445) for the first ``n_envs``
        environments, up to ``capacity`` rows per ship and episode (reset() restarts the log)."""
        n_envs = int(min(n_envs, self.num_envs))
        self.log_f64 = torch.zeros((2 * n_envs, int(capacity), len(L.LOG_COLS)), dtype=torch.float64, device=self._device)
        self.log_count = torch.zeros((2 * n_envs,), dtype=torch.int32, device=self._device)
        L.check(L.load().shipenv_set_trajectory_log(self._handle, self.log_f64.data_ptr(), self.log_count.data_ptr(),
                                                     n_envs, int(capacity)))
        self._log_envs = n_envs

    def disable_trajectory_log(self):
        torch.cuda.synchronize(self._device)
        L.check(L.load().shipenv_set_trajectory_log(self._handle, None, None, 0, 0))
        self._log_envs = 0

    def trajectory(self, role: int, env: int = 0) -> np.ndarray:
        """[rows, len(LOG_COLS)] raw log of one ship: columns ``_lib.LOG_COLS`` in SI units, values as logged by
        the reference (before the integration of the step)."""
        if not getattr(self, "_log_envs", 0) or env >= self._log_envs:
            raise RuntimeError("trajectory logging is not enabled for this environment (enable_trajectory_log)")
        i = 2 * env + role
        n = min(int(self.log_count[i].item()), self.log_f64.shape[1])
        return self.log_f64[i, :n].cpu().numpy()

    def simulation_results(self, role: int, env: int = 0) -> dict:
        """The log of one ship under the reference's ``simulation_results`` keys and units: the 12 state / controller
        columns straight from the device log (ship_model.py:418-429) and, for ShipModelAST, the 15 machinery
        bookkeeping columns of rl_env ship_model.py:911-937 (load split, powers, fuel rates and totals, motor
        torque), which are functions of the logged load fraction and shaft speed and are derived on the host
        (sim/ship_engine.py:bookkeeping_columns).  Same 27 keys, same order as the reference."""
        t = self.trajectory(role, env)
        c = {n: t[:, i] for i, n in enumerate(L.LOG_COLS)}
        deg = 180 / np.pi
        out = {
            'time [s]': c["time"], 'north position [m]': c["north"], 'east position [m]': c["east"],
            'yaw angle [deg]': c["yaw"] * deg, 'rudder angle [deg]': c["rudder"] * deg,
            'forward speed [m/s]': c["u"], 'sideways speed [m/s]': c["v"], 'yaw rate [deg/sec]': c["r"] * deg,
        }
        p = self._params.ship[role]
        if p.model_kind == L.MODEL_DETAILED:
            out['propeller shaft speed [rpm]'] = c["omega"] * 30 / np.pi
            # a row that repeats its predecessor except for the clock is a stopped ship's store_last_simulation_data
            body = np.delete(t, L.LOG_COLS.index("time"), axis=1)
            new_row = np.ones(len(t), dtype=bool)
            new_row[1:] = np.any(body[1:] != body[:-1], axis=1)
            mm = self.assets[role].ship_model.ship_machinery_model
            out.update(mm.bookkeeping_columns(c["cmd"], c["omega"], p.dt_shaft, new_row))
            out['thrust force [kN]'] = p.thrust_coeff * c["omega"] * np.abs(c["omega"]) / 1000      # ship_engine.py:411-414
        elif p.model_kind == L.MODEL_SIMPLIFIED:
            out['commanded load fraction [-]'] = c["cmd"]
            out['thrust force [kN]'] = c["omega"] / 1000          # the machinery state row is the thrust force [N]
        else:
            out['thrust force [kN]'] = c["cmd"]          # the reference logs newtons under this key (ship_model.py:427)
        out['cross track error [m]'] = c["e_ct"]
        out['heading error [deg]'] = c["e_psi"]          # radians in the reference too (controllers.py:396-397)
        return {k: v.tolist() for k, v in out.items()}

    def do_normalize_action(self, a_real):       # env.py:186-190
        return 2.0 * (a_real - self.action_space.low) / (self.action_space.high - self.action_space.low) - 1.0

    def do_denormalize_action(self, a_norm):     # env.py:192-196
        return (a_norm + 1.0) / 2.0 * (self.action_space.high - self.action_space.low) + self.action_space.low

    def seed(self, seed=None):
        self.np_random = np.random.default_rng(seed)

    # -- reference API ----------------------------------------------------------------------------
    def reset(self, action=None, mask: Optional[torch.Tensor] = None):
        """env.reset() (rl_env env.py:238-295): re-initialise (masked) environments and run init_step().
        Returns the construction-time ``initial_states`` like the reference."""
        lib = L.load()
        if not self._post_reset:
            self._post_reset = True
            if self._params.ship[0].model_kind == L.MODEL_DETAILED:
                # the first reset() leaves the machinery integrator at dt = 0.01 (ship_engine.py:331-333)
                self._params = pack_params(self.assets, self.map, self.args, self.ENV_KIND, True, self.math_mode)
                L.check(lib.shipenv_set_params(self._handle, C.byref(self._params)))
        mask_ptr = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self._device).to(torch.uint8).contiguous()
            if m.numel() != self.num_envs:
                raise ValueError("mask must have num_envs entries")
            mask_ptr = m.data_ptr()
        init_ptr = self._init_states.data_ptr() if self._init_states is not None else None
        L.check(lib.shipenv_reset(self._handle, mask_ptr, init_ptr, self._stream_ptr()))
        if self.batched:
            return self.obs_buf
        return self.initial_states

    def init_step(self, action=None):
        """env.init_step() (rl_env env.py:297-342)."""
        L.check(L.load().shipenv_init_step(self._handle, self._stream_ptr()))

    def _step(self, k: int = 1):
        """k x env._step() (rl_env env.py:563-622).  Returns what the reference's _step() returns after
        the last one."""
        L.check(L.load().shipenv_substeps(self._handle, int(k), self._stream_ptr()))
        return self._pack_step_result(with_reward=self.ENV_KIND == L.ENV_RL)

    def step(self, action):
        """env.step(action) (rl_env env.py:624-773): ``action`` holds un-normalised scoping angles [rad]
        (normalised to [-1, 1] when ``args.normalize_action``), one per environment."""
        if self.ENV_KIND == L.ENV_COLAV_NONIW and not self._params.obs_sampled_route:
            # run_colav/env.py:602: obs_ship_uses_scoping_angle() calls auto_pilot.update_route()
            raise AttributeError("'HeadingByRouteController' object has no attribute 'update_route'")
        if getattr(self.args, "normalize_action", False) and action is not None:
            if isinstance(action, torch.Tensor):      # stay on the device: same affine map as env.py:192-196
                lo = torch.as_tensor(self.action_space.low, device=action.device, dtype=action.dtype)
                hi = torch.as_tensor(self.action_space.high, device=action.device, dtype=action.dtype)
                action = (action + 1.0) / 2.0 * (hi - lo) + lo
            else:
                action = self.do_denormalize_action(action)
        a = self._actions_tensor(action)
        L.check(L.load().shipenv_step(self._handle, a.data_ptr(), self._stream_ptr()))
        return self._pack_step_result(with_reward=self.ENV_KIND == L.ENV_RL, check_unbound=True)

    def _pack_step_result(self, with_reward: bool, check_unbound: bool = False):
        if self.batched:
            done = (self.info_buf & L.INFO_DONE) != 0
            info = self._info_tensors()
            if with_reward:
                return self.obs_buf, self.reward_buf, done, info
            return self.obs_buf, done, info
        bits = int(self.info_buf[0].item())
        if check_unbound and bits & L.INFO_UNBOUND:
            # same failure as the reference (env.py:700-773 returns an unbound local)
            raise UnboundLocalError("cannot access local variable 'next_observations': step() was called after "
                                    "the sampling budget was exhausted and the obstacle ship is inside a radius "
                                    "of acceptance")
        obs = self._obs_numpy()
        done = bool(bits & L.INFO_DONE)
        info = self._info_dict_scalar(bits)
        if with_reward:
            return obs, float(self.reward_buf[0].item()), done, info
        return obs, done, info

    def ship_rollout(self, k: int):
        """k iterations of the bare per-ship loop (no env logic) for every ship."""
        L.check(L.load().shipenv_ship_rollout(self._handle, int(k), self._stream_ptr()))

    # -- host-buffer path (numpy in / numpy out through the C ABI's *_host entry points) -----------
    def _host_arrays(self):
        """The env's own host arrays of the host-buffer path: the four outputs back to back in one page-locked block
        (mirrors the device layout, so one copy fetches them) and an actions array.  Registered once through
        shipenv_register_host -- the env owns them for its whole life, so the registration cannot outlive the memory;
        close() releases them."""
        if not hasattr(self, "_host_bufs"):
            B = self.num_envs
            block = np.empty(48 * B, np.uint8)
            bufs = dict(actions=np.empty(B, np.float64), block=block,
                        obs=block[:32 * B].view(np.float32).reshape(B, 8), reward=block[32 * B:40 * B].view(np.float64),
                        info=block[40 * B:44 * B].view(np.int32), nsub=block[44 * B:48 * B].view(np.int32))
            lib = L.load()
            for key in ("actions", "block"):
                a = bufs[key]
                if a.nbytes >= (16 << 10):          # small buffers: staging is as fast as a direct copy
                    L.check(lib.shipenv_register_host(self._handle, a.ctypes.data, a.nbytes))
            self._host_bufs = bufs
        return self._host_bufs

    def register_host_actions(self, actions: np.ndarray):
        """Page-lock a caller-owned float64 array of action rows (shape [..., num_envs]) that will be passed to
        step_host() row by row and that outlives this env's use of it: step_host() then copies straight from the
        row instead of staging it.  Call unregister_host_actions() before the array is freed or reallocated."""
        a = np.ascontiguousarray(actions)
        if a is not actions or a.dtype != np.float64:
            raise ValueError("actions must be a C-contiguous float64 array")
        L.check(L.load().shipenv_register_host(self._handle, a.ctypes.data, a.nbytes))
        self._host_actions = (a.ctypes.data, a.nbytes, a)

    def unregister_host_actions(self):
        if getattr(self, "_host_actions", None) is not None:
            L.check(L.load().shipenv_unregister_host(self._handle, self._host_actions[0]))
            self._host_actions = None

    def step_host(self, actions: np.ndarray):
        """step(action) through host buffers: numpy actions in, numpy (obs, reward, info word, substeps) out.  The
        returned arrays are the env's own buffers and are overwritten by the next call."""
        B = self.num_envs
        hb = self._host_arrays()
        a = np.asarray(actions, dtype=np.float64).reshape(-1)
        if a.size != B:
            raise ValueError(f"expected {B} actions")
        reg = getattr(self, "_host_actions", None)
        if reg is not None and a.flags.c_contiguous and reg[0] <= a.ctypes.data and a.ctypes.data + a.nbytes <= reg[0] + reg[1]:
            src = a                          # a row of the array the caller registered: no staging copy
        else:
            np.copyto(hb["actions"], a)      # one persistent page-locked array instead of whatever the caller passed
            src = hb["actions"]
        L.check(L.load().shipenv_step_host(self._handle, src.ctypes.data, hb["obs"].ctypes.data,
                                           hb["reward"].ctypes.data, hb["info"].ctypes.data, hb["nsub"].ctypes.data))
        return hb["obs"], hb["reward"], hb["info"], hb["nsub"]

    def reset_host(self, mask: Optional[np.ndarray] = None):
        if not self._post_reset:
            self.reset()
        obs = self._host_arrays()["obs"]      # the array step_host() returns too: one mapped observation buffer
        mptr = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            mptr = mask.ctypes.data
        L.check(L.load().shipenv_reset_host(self._handle, mptr, obs.ctypes.data))
        return obs


class MultiShipRLEnv(BatchedShipEnv):
    """rl_env/ship_in_transit/env.py:41 -- ShipModelAST pair, AST reward, step() returns
    (next_observations, accumulated_rewards, combined_done, env_info)."""
    ENV_KIND = L.ENV_RL

    def _make_action_space(self, args):
        if getattr(args, "normalize_action", False):
            return Box(low=np.array([-1.0], dtype=np.float32), high=np.array([1.0], dtype=np.float32), dtype=np.float32)
        return Box(low=np.array([-np.deg2rad(30)], dtype=np.float32), high=np.array([np.deg2rad(30)], dtype=np.float32),
                   dtype=np.float32)


class MultiShipEnv(BatchedShipEnv):
    """run_colav/env.py:810 -- SimpleShipModel pair with intermediate-waypoint sampling; step() returns
    (next_observations, combined_done, env_info), no reward."""
    ENV_KIND = L.ENV_COLAV_IW


class MultiShipNonIWEnv(BatchedShipEnv):
    """run_colav/env.py:37 -- fixed routes, driven by init_step() + _step(); 6-entry observation."""
    ENV_KIND = L.ENV_COLAV_NONIW
    OBS_DIM = 6
