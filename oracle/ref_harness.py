"""Harness that imports the UNMODIFIED reference simulator from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.  It exists to
(1) generate the golden vectors committed under tests/golden/ (tests/golden/make_golden.py) and
(2) cross-check the C restatement in oracle/shipsim_oracle.c while /root/reference is present
(this container only -- /root/reference does not exist on the GPU box).

The reference needs four packages that are not installed here:
matplotlib, gymnasium, shapely, gtimer.  None of them does arithmetic on the simulator step
path except shapely, so they are replaced by stub modules:

* matplotlib{,.pyplot,.patches}: empty modules (imported at module top, never used on the path).
* gymnasium: ``Env`` base class, ``spaces.Box`` (low/high/shape/dtype), ``utils.seeding``.
* shapely.geometry: ``Polygon`` / ``Point`` stand-in following Shapely's documented semantics --
  ``Polygon.contains(Point)`` is the strict interior test (boundary -> False, even-odd crossing
  rule), ``polygon.exterior.distance(point)`` is the minimum Euclidean distance to the closed
  ring's segments.  The reference has no test that pins results at this boundary (SURVEY.md
  section 4), so geometry parity is "unpinned" against real GEOS; it differs from GEOS only for
  points within rounding distance of an edge.

The config builders below restate the *values* hard-coded in the reference's scripts
(run_colav/run_simplified_model.py:55-211, run_colav/run_simplified_IW_model.py:55-211,
run/env_setup.py:32-253) because those scripts cannot be imported (they execute at import time and
need the git-ignored ``test_beds`` package).
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """/root/reference in the build container; on the GPU box (where it does not exist) the byte-for-byte copy of its
    simulator packages that oracle/build_ref.py puts into oracle/_ref/ (git-ignored, travels with the snapshot)."""
    env = os.environ.get("AST_SAC_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/run_colav"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "run_colav"))


# --------------------------------------------------------------------------------------------
# stub modules
# --------------------------------------------------------------------------------------------
class _Ring:
    def __init__(self, xs, ys):
        self.xs = xs
        self.ys = ys
        self.coords = list(zip(xs + xs[:1], ys + ys[:1]))

    def distance(self, pt):
        px, py = pt.x, pt.y
        best = math.inf
        n = len(self.xs)
        for i in range(n):
            ax, ay = self.xs[i], self.ys[i]
            bx, by = self.xs[(i + 1) % n], self.ys[(i + 1) % n]
            dx, dy = bx - ax, by - ay
            l2 = dx * dx + dy * dy
            if l2 == 0.0:
                t = 0.0
            else:
                t = ((px - ax) * dx + (py - ay) * dy) / l2
                t = max(0.0, min(1.0, t))
            cx, cy = ax + t * dx, ay + t * dy
            d = math.sqrt((px - cx) * (px - cx) + (py - cy) * (py - cy))
            if d < best:
                best = d
        return best


class _Point:
    def __init__(self, x, y):
        self.x = float(x)
        self.y = float(y)


class _Polygon:
    def __init__(self, verts):
        self.xs = [float(v[0]) for v in verts]
        self.ys = [float(v[1]) for v in verts]
        self.exterior = _Ring(self.xs, self.ys)

    def contains(self, pt):
        # even-odd crossing rule; points on an edge are not "contained" up to rounding
        x, y = pt.x, pt.y
        n = len(self.xs)
        inside = False
        j = n - 1
        for i in range(n):
            xi, yi = self.xs[i], self.ys[i]
            xj, yj = self.xs[j], self.ys[j]
            if (yi > y) != (yj > y):
                if x < (xj - xi) * (y - yi) / (yj - yi) + xi:
                    inside = not inside
            j = i
        return inside


class _Box:
    def __init__(self, low, high, dtype=np.float32, shape=None):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.dtype = np.dtype(dtype)
        self.shape = self.low.shape


class _Env:
    def __init__(self, *a, **k):
        pass


def install_stubs() -> None:
    """Put the stub modules in sys.modules and /root/reference on sys.path (idempotent)."""
    if "shapely" not in sys.modules:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]

        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        gutils = types.ModuleType("gymnasium.utils")
        seeding = types.ModuleType("gymnasium.utils.seeding")
        seeding.np_random = lambda seed=None: (np.random.default_rng(seed), seed)
        gutils.seeding = seeding
        spaces.Box = _Box
        gym.Env = _Env
        gym.spaces = spaces
        gym.utils = gutils
        sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces,
                            "gymnasium.utils": gutils, "gymnasium.utils.seeding": seeding})

        shp = types.ModuleType("shapely")
        geom = types.ModuleType("shapely.geometry")
        geom.Polygon = _Polygon
        geom.Point = _Point
        shp.geometry = geom
        sys.modules.update({"shapely": shp, "shapely.geometry": geom})

        gt = types.ModuleType("gtimer")
        gt.stamp = lambda *a, **k: None
        gt.blank_stamp = lambda *a, **k: None
        sys.modules.setdefault("gtimer", gt)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


# --------------------------------------------------------------------------------------------
# shared constants of the reference's scripts
# --------------------------------------------------------------------------------------------
MAP_DATA = [
    [(0, 10000), (10000, 10000), (9200, 9000), (7600, 8500), (6700, 7300), (4900, 6500), (4300, 5400),
     (4700, 4500), (6000, 4000), (5800, 3600), (4200, 3200), (3200, 4100), (2000, 4500), (1000, 4000),
     (900, 3500), (500, 2600), (0, 2350)],
    [(10000, 0), (11500, 750), (12000, 2000), (11700, 3000), (11000, 3600), (11250, 4250), (12300, 4000),
     (13000, 3800), (14000, 3000), (14500, 2300), (15000, 1700), (16000, 800), (17500, 0)],
    [(15500, 10000), (16000, 9000), (18000, 8000), (19000, 7500), (20000, 6000), (20000, 10000)],
    [(5500, 5300), (6000, 5000), (6800, 4500), (8000, 5000), (8700, 5500), (9200, 6700), (8000, 7000),
     (6700, 6300), (6000, 6000)],
    [(15000, 5000), (14000, 5500), (12500, 5000), (14000, 4100), (16000, 2000), (15700, 3700)],
    [(11000, 2000), (10300, 3200), (9000, 1500), (10000, 1000)],
]


class Args:
    """Stand-in for the argparse namespace (run/env_args.py:8-22)."""

    def __init__(self, time_step=4, collav_mode="none", max_sampling_frequency=9,
                 radius_of_acceptance=300, lookahead_distance=1000, normalize_action=False):
        self.max_sampling_frequency = max_sampling_frequency
        self.time_step = time_step
        self.radius_of_acceptance = radius_of_acceptance
        self.lookahead_distance = lookahead_distance
        self.collav_mode = collav_mode
        self.ship_draw = False
        self.time_since_last_ship_drawing = 30
        self.normalize_action = normalize_action


SHIP_CONFIG = dict(
    coefficient_of_deadweight_to_displacement=0.7, bunkers=200000, ballast=200000,
    length_of_ship=80, width_of_ship=16,
    added_mass_coefficient_in_surge=0.4, added_mass_coefficient_in_sway=0.4,
    added_mass_coefficient_in_yaw=0.4, dead_weight_tonnage=3850000,
    mass_over_linear_friction_coefficient_in_surge=130,
    mass_over_linear_friction_coefficient_in_sway=18,
    mass_over_linear_friction_coefficient_in_yaw=90,
    nonlinear_friction_coefficient__in_surge=2400,
    nonlinear_friction_coefficient__in_sway=4000,
    nonlinear_friction_coefficient__in_yaw=400)
ENV_CONFIG = dict(current_velocity_component_from_north=-1, current_velocity_component_from_east=-1,
                  wind_speed=2, wind_direction=-np.pi / 4)
TEST_INIT = dict(initial_north_position_m=100, initial_east_position_m=100,
                 initial_yaw_angle_rad=60 * np.pi / 180, initial_forward_speed_m_per_s=4.25,
                 initial_sideways_speed_m_per_s=0, initial_yaw_rate_rad_per_s=0)
OBS_INIT = dict(initial_north_position_m=9900, initial_east_position_m=14900,
                initial_yaw_angle_rad=-135 * np.pi / 180, initial_forward_speed_m_per_s=3.5,
                initial_sideways_speed_m_per_s=0, initial_yaw_rate_rad_per_s=0)


def build_colav_assets(args: Args, obs_route="obs_ship_route.txt", test_init=None, obs_init=None,
                       sim_time=10000):
    """run_colav/run_simplified_IW_model.py:55-211 (obs_ship_route.txt) or
    run_colav/run_simplified_model.py:55-211 (obs_ship_route_nonIW.txt) -- the SimpleShipModel pair."""
    install_stubs()
    from run_colav.ship_in_transit.sub_systems.ship_model import (
        ShipConfiguration, EnvironmentConfiguration, SimulationConfiguration, SimpleShipModel)
    from run_colav.ship_in_transit.sub_systems.ship_engine import RudderConfiguration
    from run_colav.ship_in_transit.sub_systems.controllers import (
        SpeedControllerGains, HeadingControllerGains, LosParameters, ThrustFromSpeedSetPoint,
        HeadingBySampledRouteController)
    from run_colav.ship_in_transit.sub_systems.obstacle import PolygonObstacle
    from run_colav.env import ShipAssets
    from utils.paths_utils import get_data_path_run_colav

    ship_config = ShipConfiguration(**SHIP_CONFIG)
    env_config = EnvironmentConfiguration(**ENV_CONFIG)
    rudder_config = RudderConfiguration(rudder_angle_to_sway_force_coefficient=50e3,
                                        rudder_angle_to_yaw_force_coefficient=500e3,
                                        max_rudder_angle_degrees=30)
    ti = dict(TEST_INIT, **(test_init or {}))
    oi = dict(OBS_INIT, **(obs_init or {}))
    test_ship = SimpleShipModel(ship_config=ship_config, rudder_config=rudder_config,
                                environment_config=env_config,
                                simulation_config=SimulationConfiguration(
                                    integration_step=args.time_step, simulation_time=sim_time, **ti))
    obs_ship = SimpleShipModel(ship_config=ship_config, rudder_config=rudder_config,
                               environment_config=env_config,
                               simulation_config=SimulationConfiguration(
                                   integration_step=args.time_step, simulation_time=sim_time, **oi))
    map_obj = PolygonObstacle(MAP_DATA)

    def los():
        return LosParameters(radius_of_acceptance=args.radius_of_acceptance,
                             lookahead_distance=args.lookahead_distance,
                             integral_gain=0.002, integrator_windup_limit=4000)

    test_ctrl = ThrustFromSpeedSetPoint(gains=SpeedControllerGains(kp=150, ki=150, kd=75),
                                        max_thrust=np.inf, time_step=args.time_step)
    test_ap = HeadingBySampledRouteController(
        get_data_path_run_colav("own_ship_route.txt"),
        heading_controller_gains=HeadingControllerGains(kp=.5, ki=0.01, kd=84),
        los_parameters=los(), time_step=args.time_step,
        max_rudder_angle=np.deg2rad(rudder_config.max_rudder_angle_degrees), num_of_samplings=2)
    obs_ctrl = ThrustFromSpeedSetPoint(gains=SpeedControllerGains(kp=.025, ki=700.5, kd=550.5),
                                       max_thrust=np.inf, time_step=args.time_step)
    obs_ap = HeadingBySampledRouteController(
        get_data_path_run_colav(obs_route),
        heading_controller_gains=HeadingControllerGains(kp=.65, ki=0.001, kd=50),
        los_parameters=los(), time_step=args.time_step,
        max_rudder_angle=np.deg2rad(rudder_config.max_rudder_angle_degrees), num_of_samplings=2)
    test = ShipAssets(ship_model=test_ship, speed_controller=test_ctrl, auto_pilot=test_ap,
                      desired_forward_speed=4.5, integrator_term=[], time_list=[],
                      stop_flag=False, type_tag='test_ship')
    obs = ShipAssets(ship_model=obs_ship, speed_controller=obs_ctrl, auto_pilot=obs_ap,
                     desired_forward_speed=4.0, integrator_term=[], time_list=[],
                     stop_flag=False, type_tag='obs_ship')
    return [test, obs], map_obj


MACHINERY_MODES = {
    # run/env_setup.py:62-81
    "PTO": dict(main_engine_capacity=2160e3, electrical_capacity=0, shaft_generator_state='GEN'),
    "PTI": dict(main_engine_capacity=0, electrical_capacity=2 * 510e3, shaft_generator_state='MOTOR'),
    "MEC": dict(main_engine_capacity=2160e3, electrical_capacity=510e3, shaft_generator_state='OFF'),
}


def build_rl_assets(args: Args, mode="PTI", test_init=None, obs_init=None, sim_time=10000, omega_init=None):
    """run/env_setup.py:32-239 -- the ShipModelAST pair (detailed machinery).  ``omega_init`` = (test, obs) initial
    propeller shaft speeds [rad/s] when they differ from the script's 420 / 200 rpm (the one-ulp twins of
    tests/golden/make_reference_twins.py)."""
    install_stubs()
    from rl_env.ship_in_transit.env import ShipAssets
    from rl_env.ship_in_transit.sub_systems.ship_model import (
        ShipConfiguration, EnvironmentConfiguration, SimulationConfiguration, ShipModelAST)
    from rl_env.ship_in_transit.sub_systems.ship_engine import (
        MachinerySystemConfiguration, MachineryMode, MachineryModeParams, MachineryModes,
        SpecificFuelConsumptionBaudouin6M26Dot3, SpecificFuelConsumptionWartila6L26)
    from rl_env.ship_in_transit.sub_systems.LOS_guidance import LosParameters
    from rl_env.ship_in_transit.sub_systems.obstacle import PolygonObstacle
    from rl_env.ship_in_transit.sub_systems.controllers import (
        ThrottleControllerGains, HeadingControllerGains, EngineThrottleFromSpeedSetPoint,
        HeadingBySampledRouteController)
    from utils.paths_utils import get_data_path

    ship_config = ShipConfiguration(**SHIP_CONFIG)
    env_config = EnvironmentConfiguration(**ENV_CONFIG)

    def machinery():
        mso_modes = MachineryModes([MachineryMode(params=MachineryModeParams(**MACHINERY_MODES[mode]))])
        return MachinerySystemConfiguration(
            machinery_modes=mso_modes, machinery_operating_mode=0,
            linear_friction_main_engine=68, linear_friction_hybrid_shaft_generator=57,
            gear_ratio_between_main_engine_and_propeller=0.6,
            gear_ratio_between_hybrid_shaft_generator_and_propeller=0.6,
            propeller_inertia=6000, propeller_diameter=3.1,
            propeller_speed_to_torque_coefficient=7.5,
            propeller_speed_to_thrust_force_coefficient=1.7,
            hotel_load=200000, rated_speed_main_engine_rpm=1000,
            rudder_angle_to_sway_force_coefficient=50e3,
            rudder_angle_to_yaw_force_coefficient=500e3, max_rudder_angle_degrees=30,
            specific_fuel_consumption_coefficients_me=SpecificFuelConsumptionWartila6L26().fuel_consumption_coefficients(),
            specific_fuel_consumption_coefficients_dg=SpecificFuelConsumptionBaudouin6M26Dot3().fuel_consumption_coefficients())

    # NOTE: the reference shares one machinery_config (hence one MachineryMode object) between
    # both ships (run/env_setup.py:85-105,120-142); the mode object is read-only on the step path.
    machinery_config = machinery()
    ti = dict(TEST_INIT, **(test_init or {}))
    oi = dict(OBS_INIT, **(obs_init or {}))
    test_ship = ShipModelAST(ship_config=ship_config, machinery_config=machinery_config,
                             environment_config=env_config,
                             simulation_config=SimulationConfiguration(
                                 integration_step=args.time_step, simulation_time=sim_time, **ti),
                             initial_propeller_shaft_speed_rad_per_s=(omega_init[0] if omega_init else 420 * np.pi / 30))
    obs_ship = ShipModelAST(ship_config=ship_config, machinery_config=machinery_config,
                            environment_config=env_config,
                            simulation_config=SimulationConfiguration(
                                integration_step=args.time_step, simulation_time=sim_time, **oi),
                            initial_propeller_shaft_speed_rad_per_s=(omega_init[1] if omega_init else 200 * np.pi / 30))
    map_obj = PolygonObstacle(MAP_DATA)
    gains = dict(kp_ship_speed=205.25, ki_ship_speed=0.0525, kp_shaft_speed=50, ki_shaft_speed=0.00025)

    def los():
        return LosParameters(radius_of_acceptance=args.radius_of_acceptance,
                             lookahead_distance=args.lookahead_distance,
                             integral_gain=0.002, integrator_windup_limit=4000)

    def throttle_ctrl(ship):
        return EngineThrottleFromSpeedSetPoint(
            gains=ThrottleControllerGains(**gains),
            max_shaft_speed=ship.ship_machinery_model.shaft_speed_max,
            time_step=args.time_step, initial_shaft_speed_integral_error=114)

    def autopilot(route):
        return HeadingBySampledRouteController(
            get_data_path(route),
            heading_controller_gains=HeadingControllerGains(kp=1.65, kd=75, ki=0.001),
            los_parameters=los(), time_step=args.time_step,
            max_rudder_angle=machinery_config.max_rudder_angle_degrees * np.pi / 180,
            num_of_samplings=2)

    test = ShipAssets(ship_model=test_ship, throttle_controller=throttle_ctrl(test_ship),
                      auto_pilot=autopilot('test_ship_route.txt'), desired_forward_speed=4.5,
                      integrator_term=[], time_list=[], stop_flag=False, type_tag='test_ship')
    obs = ShipAssets(ship_model=obs_ship, throttle_controller=throttle_ctrl(obs_ship),
                     auto_pilot=autopilot('obs_ship_route.txt'), desired_forward_speed=4.0,
                     integrator_term=[], time_list=[], stop_flag=False, type_tag='obs_ship')
    return [test, obs], map_obj


def make_colav_iw_env(args: Args, **kw):
    assets, map_obj = build_colav_assets(args, obs_route="obs_ship_route.txt", **kw)
    from run_colav.env import MultiShipEnv
    return MultiShipEnv(assets=assets, map=map_obj, args=args), assets


def make_colav_noniw_env(args: Args, obs_route="obs_ship_route_nonIW.txt", **kw):
    assets, map_obj = build_colav_assets(args, obs_route=obs_route, **kw)
    from run_colav.env import MultiShipNonIWEnv
    return MultiShipNonIWEnv(assets=assets, map=map_obj, args=args), assets


def make_rl_env(args: Args, **kw):
    assets, map_obj = build_rl_assets(args, **kw)
    from rl_env.ship_in_transit.env import MultiShipRLEnv
    return MultiShipRLEnv(assets=assets, map=map_obj, args=args), assets


# --------------------------------------------------------------------------------------------
# timing of the unmodified reference (bench.py cpu_baseline: "reference_python_1core")
# --------------------------------------------------------------------------------------------
def time_reference(workload: str = "colav_iw", collav: str = "none", min_steps: int = 2000, warmup_steps: int = 200,
                   seed: int = 0):
    """env-steps/s of the UNMODIFIED reference env on one host core (the reference is single-process Python):
    episodes of reset() + 9 step(action) with scoping angles ~ U(-pi/6, pi/6), `warmup_steps` simulator steps
    untimed, then whole step() calls until at least `min_steps` simulator steps are timed.
    Returns (rate, steps, seconds).  One env-step = one _step() = one row of the ship's simulation log."""
    import time
    rng = np.random.default_rng(seed)
    args = Args(time_step=4, collav_mode=collav)
    env, assets = (make_rl_env(args) if workload == "rl" else make_colav_iw_env(args))
    log = assets[1].ship_model.simulation_results

    def rows():
        return len(log['time [s]']) if 'time [s]' in log else 0

    timed_steps, timed_s, warmed = 0, 0.0, 0
    while timed_steps < min_steps:
        env.reset()
        log = assets[1].ship_model.simulation_results
        for _ in range(9):
            a = np.array([rng.uniform(-np.pi / 6, np.pi / 6)], dtype=np.float64)
            n0 = rows()
            t0 = time.perf_counter()
            out = env.step(a)
            dt = time.perf_counter() - t0
            n = rows() - n0
            if warmed < warmup_steps:
                warmed += n
            else:
                timed_steps += n
                timed_s += dt
            if out[-2] or timed_steps >= min_steps:
                break
    return timed_steps / timed_s, timed_steps, timed_s
