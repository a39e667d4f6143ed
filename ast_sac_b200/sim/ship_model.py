"""Ship model configuration holders.  Mirrors */ship_in_transit/sub_systems/ship_model.py of the
reference: ShipConfiguration/EnvironmentConfiguration/SimulationConfiguration (ship_model.py:20-53),
BaseShipModel's derived constants (:70-132), SimpleShipModel (:322-349), ShipModelAST (rl_env
ship_model.py:803-832).  The derived constants are computed with the reference's own expressions so
their bits are identical; the dynamics themselves run on the GPU."""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from .ship_engine import MachinerySystemConfiguration, RudderConfiguration, ShipMachineryModel, _Integrator


class ShipConfiguration(NamedTuple):
    dead_weight_tonnage: float
    coefficient_of_deadweight_to_displacement: float
    bunkers: float
    ballast: float
    length_of_ship: float
    width_of_ship: float
    added_mass_coefficient_in_surge: float
    added_mass_coefficient_in_sway: float
    added_mass_coefficient_in_yaw: float
    mass_over_linear_friction_coefficient_in_surge: float
    mass_over_linear_friction_coefficient_in_sway: float
    mass_over_linear_friction_coefficient_in_yaw: float
    nonlinear_friction_coefficient__in_surge: float
    nonlinear_friction_coefficient__in_sway: float
    nonlinear_friction_coefficient__in_yaw: float


class EnvironmentConfiguration(NamedTuple):
    current_velocity_component_from_north: float
    current_velocity_component_from_east: float
    wind_speed: float
    wind_direction: float


class SimulationConfiguration(NamedTuple):
    initial_north_position_m: float
    initial_east_position_m: float
    initial_yaw_angle_rad: float
    initial_forward_speed_m_per_s: float
    initial_sideways_speed_m_per_s: float
    initial_yaw_rate_rad_per_s: float
    integration_step: float
    simulation_time: float


class BaseShipModel:
    _STATE_ROWS = {"north": 0, "east": 1, "yaw_angle": 2, "forward_speed": 3, "sideways_speed": 4, "yaw_rate": 5}

    def __init__(self, ship_config: ShipConfiguration, simulation_config: SimulationConfiguration,
                 environment_config: EnvironmentConfiguration):
        self.ship_config = ship_config
        self.simulation_config = simulation_config
        self.environment_config = environment_config
        payload = 0.9 * (ship_config.dead_weight_tonnage - ship_config.bunkers)
        lsw = ship_config.dead_weight_tonnage / ship_config.coefficient_of_deadweight_to_displacement \
            - ship_config.dead_weight_tonnage
        self.mass = lsw + payload + ship_config.bunkers + ship_config.ballast
        self.l_ship = ship_config.length_of_ship
        self.w_ship = ship_config.width_of_ship
        self.x_g = 0
        self.i_z = self.mass * (self.l_ship ** 2 + self.w_ship ** 2) / 12
        self.x_du = self.mass * ship_config.added_mass_coefficient_in_surge
        self.y_dv = self.mass * ship_config.added_mass_coefficient_in_sway
        self.n_dr = self.i_z * ship_config.added_mass_coefficient_in_yaw
        self.t_surge = ship_config.mass_over_linear_friction_coefficient_in_surge
        self.t_sway = ship_config.mass_over_linear_friction_coefficient_in_sway
        self.t_yaw = ship_config.mass_over_linear_friction_coefficient_in_yaw
        self.ku = ship_config.nonlinear_friction_coefficient__in_surge
        self.kv = ship_config.nonlinear_friction_coefficient__in_sway
        self.kr = ship_config.nonlinear_friction_coefficient__in_yaw
        self.vel_c = np.array([environment_config.current_velocity_component_from_north,
                               environment_config.current_velocity_component_from_east, 0.0])
        self.wind_dir = environment_config.wind_direction
        self.wind_speed = environment_config.wind_speed
        self._init_state = dict(
            north=np.float64(simulation_config.initial_north_position_m),
            east=np.float64(simulation_config.initial_east_position_m),
            yaw_angle=np.float64(simulation_config.initial_yaw_angle_rad),
            forward_speed=np.float64(simulation_config.initial_forward_speed_m_per_s),
            sideways_speed=np.float64(simulation_config.initial_sideways_speed_m_per_s),
            yaw_rate=np.float64(simulation_config.initial_yaw_rate_rad_per_s))
        self.int = _Integrator(dt=simulation_config.integration_step, sim_time=simulation_config.simulation_time)
        self.rho_a = 1.2
        self.h_f = 8.0
        self.h_s = 8.0
        self.proj_area_f = self.w_ship * self.h_f
        self.proj_area_l = self.l_ship * self.h_s
        self.cx, self.cy, self.cn = 0.5, 0.7, 0.08
        self._binding = None          # (env, role) once the asset is handed to a batched env

    def __getattr__(self, name):
        # live state of environment 0 (device read) once bound, else the configured initial value
        rows = BaseShipModel._STATE_ROWS
        if name in rows:
            binding = self.__dict__.get("_binding")
            if binding is not None:
                env, role = binding
                return env.read_ship_state(role)[rows[name]]
            return self.__dict__["_init_state"][name]
        if name == "simulation_results":
            # the reference's per-step log (ship_model.py:418-429 / rl_env :903-942) of environment 0, available
            # after env.enable_trajectory_log()
            binding = self.__dict__.get("_binding")
            if binding is not None:
                env, role = binding
                return env.simulation_results(role)
        raise AttributeError(name)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_binding"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)


class SimpleShipModel(BaseShipModel):
    def __init__(self, ship_config: ShipConfiguration, simulation_config: SimulationConfiguration,
                 environment_config: EnvironmentConfiguration, rudder_config: RudderConfiguration):
        super().__init__(ship_config, simulation_config, environment_config)
        self.rudder_config = RudderConfiguration(
            rudder_angle_to_sway_force_coefficient=rudder_config.rudder_angle_to_sway_force_coefficient,
            rudder_angle_to_yaw_force_coefficient=rudder_config.rudder_angle_to_yaw_force_coefficient,
            max_rudder_angle_degrees=rudder_config.max_rudder_angle_degrees)


class ShipModelAST(BaseShipModel):
    def __init__(self, ship_config: ShipConfiguration, simulation_config: SimulationConfiguration,
                 environment_config: EnvironmentConfiguration, machinery_config: MachinerySystemConfiguration,
                 initial_propeller_shaft_speed_rad_per_s):
        super().__init__(ship_config, simulation_config, environment_config)
        self.ship_machinery_model = ShipMachineryModel(
            machinery_config=machinery_config,
            initial_propeller_shaft_speed_rad_per_sec=initial_propeller_shaft_speed_rad_per_s,
            time_step=self.int.dt)


class ShipModelSimplifiedPropulsion(BaseShipModel):
    """Hull driven by SimplifiedMachineryModel (thrust-force state T).  The reference ships the machinery model
    (ship_engine.py:484-519) and its throttle controller (rl_env controllers.py:212-232) but no ship model class
    that uses them; this one wires them the way ShipModelAST wires the detailed machinery
    (rl_env ship_model.py:882-901): the current thrust state feeds the kinetics, then hull and thrust state are
    integrated."""

    def __init__(self, ship_config: ShipConfiguration, simulation_config: SimulationConfiguration,
                 environment_config: EnvironmentConfiguration, machinery_config, initial_thrust_force: float = 0.0):
        super().__init__(ship_config, simulation_config, environment_config)
        from .ship_engine import SimplifiedMachineryModel
        self.ship_machinery_model = SimplifiedMachineryModel(machinery_config=machinery_config, time_step=self.int.dt,
                                                             initial_thrust_force=initial_thrust_force)


ShipModel = ShipModelAST      # identical dynamics (rl_env ship_model.py:664 vs :803); only the log keys differ
