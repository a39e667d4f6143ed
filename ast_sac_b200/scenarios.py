"""The reference's hard-coded scenarios, rebuilt on the batched environments.

* ``prepare_multiship_rl_env(args, ...)``  -- run/env_setup.py:17-253 (ShipModelAST pair, PTI/PTO/MEC)
* ``prepare_colav_env(args, iw=True/False)`` -- run_colav/run_simplified_IW_model.py:54-233 and
  run_colav/run_simplified_model.py:54-205 (SimpleShipModel pair)
* ``get_env_args(...)``                     -- run/env_args.py:3-27 defaults as a namespace

Only the numbers are shared with the reference; the objects are this package's parameter holders.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import List

import numpy as np

from .env import MultiShipEnv, MultiShipNonIWEnv, MultiShipRLEnv, ShipAssets
from .sim.controllers import (EngineThrottleFromSpeedSetPoint, HeadingBySampledRouteController,
                              HeadingControllerGains, LosParameters, SpeedControllerGains,
                              ThrottleControllerGains, ThrustFromSpeedSetPoint)
from .sim.obstacle import PolygonObstacle
from .sim.ship_engine import (MachineryMode, MachineryModeParams, MachineryModes, MachinerySystemConfiguration,
                              RudderConfiguration, SpecificFuelConsumptionBaudouin6M26Dot3,
                              SpecificFuelConsumptionWartila6L26)
from .sim.ship_model import (EnvironmentConfiguration, ShipConfiguration, ShipModelAST, SimpleShipModel,
                             SimulationConfiguration)

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def get_data_path(filename: str) -> str:
    """utils/paths_utils.py:6-10 equivalent for the route files shipped with this package."""
    return os.path.join(DATA_DIR, filename)


def get_env_args(max_sampling_frequency=9, time_step=4, radius_of_acceptance=300, lookahead_distance=1000,
                 collav_mode='none', ship_draw=False, time_since_last_ship_drawing=30, normalize_action=False):
    """run/env_args.py:8-22.  The reference's default collav_mode is 'sbmpc' (run/ast-sac_runner.py:35); the parity and
    bench scenarios here default to 'none' and name the mode explicitly."""
    return SimpleNamespace(max_sampling_frequency=max_sampling_frequency, time_step=time_step,
                           radius_of_acceptance=radius_of_acceptance, lookahead_distance=lookahead_distance,
                           collav_mode=collav_mode, ship_draw=ship_draw,
                           time_since_last_ship_drawing=time_since_last_ship_drawing,
                           normalize_action=normalize_action)


MAP_DATA = [  # run/env_setup.py:145-152, (east, north) vertices of the six islands
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

MACHINERY_MODES = {  # run/env_setup.py:62-81
    "PTO": MachineryModeParams(main_engine_capacity=2160e3, electrical_capacity=0, shaft_generator_state='GEN'),
    "PTI": MachineryModeParams(main_engine_capacity=0, electrical_capacity=2 * 510e3, shaft_generator_state='MOTOR'),
    "MEC": MachineryModeParams(main_engine_capacity=2160e3, electrical_capacity=510e3, shaft_generator_state='OFF'),
}


def _ship_config():
    return ShipConfiguration(
        coefficient_of_deadweight_to_displacement=0.7, bunkers=200000, ballast=200000, length_of_ship=80,
        width_of_ship=16, added_mass_coefficient_in_surge=0.4, added_mass_coefficient_in_sway=0.4,
        added_mass_coefficient_in_yaw=0.4, dead_weight_tonnage=3850000,
        mass_over_linear_friction_coefficient_in_surge=130, mass_over_linear_friction_coefficient_in_sway=18,
        mass_over_linear_friction_coefficient_in_yaw=90, nonlinear_friction_coefficient__in_surge=2400,
        nonlinear_friction_coefficient__in_sway=4000, nonlinear_friction_coefficient__in_yaw=400)


def _env_config():
    return EnvironmentConfiguration(current_velocity_component_from_north=-1, current_velocity_component_from_east=-1,
                                    wind_speed=2, wind_direction=-np.pi / 4)


def _sim_config(args, who: str, override=None, sim_time=10000):
    base = dict(initial_north_position_m=100, initial_east_position_m=100, initial_yaw_angle_rad=60 * np.pi / 180,
                initial_forward_speed_m_per_s=4.25, initial_sideways_speed_m_per_s=0, initial_yaw_rate_rad_per_s=0) \
        if who == "test" else \
        dict(initial_north_position_m=9900, initial_east_position_m=14900, initial_yaw_angle_rad=-135 * np.pi / 180,
             initial_forward_speed_m_per_s=3.5, initial_sideways_speed_m_per_s=0, initial_yaw_rate_rad_per_s=0)
    base.update(override or {})
    return SimulationConfiguration(integration_step=args.time_step, simulation_time=sim_time, **base)


def _los(args):
    return LosParameters(radius_of_acceptance=args.radius_of_acceptance, lookahead_distance=args.lookahead_distance,
                         integral_gain=0.002, integrator_windup_limit=4000)


def build_rl_assets(args, mode="PTI", test_init=None, obs_init=None, sim_time=10000):
    """run/env_setup.py:32-239."""
    ship_config, env_config = _ship_config(), _env_config()
    mso_modes = MachineryModes([MachineryMode(params=MACHINERY_MODES[mode])])
    machinery_config = MachinerySystemConfiguration(
        machinery_modes=mso_modes, machinery_operating_mode=0, linear_friction_main_engine=68,
        linear_friction_hybrid_shaft_generator=57, gear_ratio_between_main_engine_and_propeller=0.6,
        gear_ratio_between_hybrid_shaft_generator_and_propeller=0.6, propeller_inertia=6000, propeller_diameter=3.1,
        propeller_speed_to_torque_coefficient=7.5, propeller_speed_to_thrust_force_coefficient=1.7,
        hotel_load=200000, rated_speed_main_engine_rpm=1000, rudder_angle_to_sway_force_coefficient=50e3,
        rudder_angle_to_yaw_force_coefficient=500e3, max_rudder_angle_degrees=30,
        specific_fuel_consumption_coefficients_me=SpecificFuelConsumptionWartila6L26().fuel_consumption_coefficients(),
        specific_fuel_consumption_coefficients_dg=SpecificFuelConsumptionBaudouin6M26Dot3().fuel_consumption_coefficients())
    test_ship = ShipModelAST(ship_config=ship_config, machinery_config=machinery_config, environment_config=env_config,
                             simulation_config=_sim_config(args, "test", test_init, sim_time),
                             initial_propeller_shaft_speed_rad_per_s=420 * np.pi / 30)
    obs_ship = ShipModelAST(ship_config=ship_config, machinery_config=machinery_config, environment_config=env_config,
                            simulation_config=_sim_config(args, "obs", obs_init, sim_time),
                            initial_propeller_shaft_speed_rad_per_s=200 * np.pi / 30)

    def throttle(ship):
        return EngineThrottleFromSpeedSetPoint(
            gains=ThrottleControllerGains(kp_ship_speed=205.25, ki_ship_speed=0.0525, kp_shaft_speed=50,
                                          ki_shaft_speed=0.00025),
            max_shaft_speed=ship.ship_machinery_model.shaft_speed_max, time_step=args.time_step,
            initial_shaft_speed_integral_error=114)

    def autopilot(route):
        return HeadingBySampledRouteController(
            get_data_path(route), heading_controller_gains=HeadingControllerGains(kp=1.65, kd=75, ki=0.001),
            los_parameters=_los(args), time_step=args.time_step,
            max_rudder_angle=machinery_config.max_rudder_angle_degrees * np.pi / 180, num_of_samplings=2)

    test = ShipAssets(ship_model=test_ship, throttle_controller=throttle(test_ship),
                      auto_pilot=autopilot('test_ship_route.txt'), desired_forward_speed=4.5, integrator_term=[],
                      time_list=[], stop_flag=False, type_tag='test_ship')
    obs = ShipAssets(ship_model=obs_ship, throttle_controller=throttle(obs_ship),
                     auto_pilot=autopilot('obs_ship_route.txt'), desired_forward_speed=4.0, integrator_term=[],
                     time_list=[], stop_flag=False, type_tag='obs_ship')
    return [test, obs], PolygonObstacle(MAP_DATA)


def build_colav_assets(args, iw=True, test_init=None, obs_init=None, sim_time=10000, obs_route=None):
    """run_colav/run_simplified_IW_model.py:55-211 (iw=True) / run_simplified_model.py:55-211 (iw=False).
    ``obs_route`` overrides the obstacle ship's route file (default: the script's own)."""
    ship_config, env_config = _ship_config(), _env_config()
    rudder_config = RudderConfiguration(rudder_angle_to_sway_force_coefficient=50e3,
                                        rudder_angle_to_yaw_force_coefficient=500e3, max_rudder_angle_degrees=30)
    test_ship = SimpleShipModel(ship_config=ship_config, rudder_config=rudder_config, environment_config=env_config,
                                simulation_config=_sim_config(args, "test", test_init, sim_time))
    obs_ship = SimpleShipModel(ship_config=ship_config, rudder_config=rudder_config, environment_config=env_config,
                               simulation_config=_sim_config(args, "obs", obs_init, sim_time))
    max_rudder = np.deg2rad(rudder_config.max_rudder_angle_degrees)
    test = ShipAssets(
        ship_model=test_ship,
        speed_controller=ThrustFromSpeedSetPoint(gains=SpeedControllerGains(kp=150, ki=150, kd=75), max_thrust=np.inf,
                                                 time_step=args.time_step),
        auto_pilot=HeadingBySampledRouteController(
            get_data_path('own_ship_route.txt'), heading_controller_gains=HeadingControllerGains(kp=.5, ki=0.01, kd=84),
            los_parameters=_los(args), time_step=args.time_step, max_rudder_angle=max_rudder, num_of_samplings=2),
        desired_forward_speed=4.5, integrator_term=[], time_list=[], stop_flag=False, type_tag='test_ship')
    obs = ShipAssets(
        ship_model=obs_ship,
        speed_controller=ThrustFromSpeedSetPoint(gains=SpeedControllerGains(kp=.025, ki=700.5, kd=550.5),
                                                 max_thrust=np.inf, time_step=args.time_step),
        auto_pilot=HeadingBySampledRouteController(
            get_data_path(obs_route or ('obs_ship_route.txt' if iw else 'obs_ship_route_nonIW.txt')),
            heading_controller_gains=HeadingControllerGains(kp=.65, ki=0.001, kd=50),
            los_parameters=_los(args), time_step=args.time_step, max_rudder_angle=max_rudder, num_of_samplings=2),
        desired_forward_speed=4.0, integrator_term=[], time_list=[], stop_flag=False, type_tag='obs_ship')
    return [test, obs], PolygonObstacle(MAP_DATA)


def build_simplified_assets(args, sim_time=10000, thrust_force_dynamic_time_constant=30.0, initial_thrust_force=0.0,
                            kp=3.0, ki=0.02):
    """The run/env_setup.py pair with the hull driven by SimplifiedMachineryModel (thrust-force state T) and
    ThrottleFromSpeedSetPointSimplifiedPropulsion instead of the detailed machinery (A8' of SURVEY.md section 8a;
    numbers of tests/golden/make_golden.py:SIMPLIFIED)."""
    from .sim.controllers import ThrottleFromSpeedSetPointSimplifiedPropulsion
    from .sim.ship_engine import SimplifiedPropulsionMachinerySystemConfiguration
    from .sim.ship_model import ShipModelSimplifiedPropulsion
    assets, m = build_rl_assets(args, sim_time=sim_time)
    out = []
    for a in assets:
        full = a.ship_model.ship_machinery_model
        cfg = SimplifiedPropulsionMachinerySystemConfiguration(
            hotel_load=full.hotel_load, machinery_modes=full.machinery_modes, machinery_operating_mode=0,
            specific_fuel_consumption_coefficients_me=None, specific_fuel_consumption_coefficients_dg=None,
            thrust_force_dynamic_time_constant=thrust_force_dynamic_time_constant,
            rudder_angle_to_sway_force_coefficient=full.c_rudder_v, rudder_angle_to_yaw_force_coefficient=full.c_rudder_r,
            max_rudder_angle_degrees=30)
        sm = a.ship_model
        model = ShipModelSimplifiedPropulsion(ship_config=sm.ship_config, simulation_config=sm.simulation_config,
                                              environment_config=sm.environment_config, machinery_config=cfg,
                                              initial_thrust_force=initial_thrust_force)
        out.append(ShipAssets(ship_model=model, auto_pilot=a.auto_pilot, desired_forward_speed=a.desired_forward_speed,
                              integrator_term=[], time_list=[], type_tag=a.type_tag, stop_flag=False,
                              throttle_controller=ThrottleFromSpeedSetPointSimplifiedPropulsion(kp=kp, ki=ki, time_step=args.time_step)))
    return out, m


def prepare_multiship_rl_env(args, num_envs=1, device=None, mode="PTI", init_states=None, math_mode=None, **kw):
    """Drop-in for run/env_setup.py:prepare_multiship_rl_env -> (env, assets)."""
    assets, map_obj = build_rl_assets(args, mode=mode, **kw)
    env = MultiShipRLEnv(assets=assets, map=map_obj, args=args, num_envs=num_envs, device=device,
                         init_states=init_states, math_mode=math_mode)
    return env, assets


def prepare_colav_env(args, iw=True, num_envs=1, device=None, init_states=None, math_mode=None, **kw):
    assets, map_obj = build_colav_assets(args, iw=iw, **kw)
    cls = MultiShipEnv if iw else MultiShipNonIWEnv
    env = cls(assets=assets, map=map_obj, args=args, num_envs=num_envs, device=device, init_states=init_states,
              math_mode=math_mode)
    return env, assets


def _machinery_state(ship_model) -> float:
    """Initial value of the machinery state row: shaft speed (ShipModelAST), thrust force (thrust-state model), 0."""
    mm = getattr(ship_model, "ship_machinery_model", None)
    if mm is None:
        return 0.0
    return float(mm.omega) if hasattr(mm, "omega") else float(mm.thrust)


def jittered_init_states(assets: List[ShipAssets], num_envs: int, pos_jitter_m: float = 100.0, seed: int = 1,
                         device="cuda"):
    """Per-environment initial states [7, 2 * num_envs]: the assets' configured initial values with a
    uniform +-pos_jitter_m offset on north/east (de-synchronises the
    data-dependent control flow between environments)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    base = torch.zeros((7, num_envs, 2), dtype=torch.float64)
    for role, a in enumerate(assets):
        sc = a.ship_model.simulation_config
        vals = [sc.initial_north_position_m, sc.initial_east_position_m, sc.initial_yaw_angle_rad,
                sc.initial_forward_speed_m_per_s, sc.initial_sideways_speed_m_per_s, sc.initial_yaw_rate_rad_per_s,
                _machinery_state(a.ship_model)]
        for i, v in enumerate(vals):
            base[i, :, role] = float(v)
    jit = (torch.rand((2, num_envs, 2), generator=g, dtype=torch.float64) * 2.0 - 1.0) * pos_jitter_m
    base[0:2] += jit
    return base.reshape(7, 2 * num_envs).to(device)
