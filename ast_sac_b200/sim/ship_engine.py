"""Machinery configuration holders.  Mirrors */ship_in_transit/sub_systems/ship_engine.py of the
reference (class and field names: ship_engine.py:17-163, derived attributes :342-370)."""
from __future__ import annotations

from typing import List, NamedTuple, Union

import numpy as np


class MachineryModeParams(NamedTuple):           # ship_engine.py:17-20
    main_engine_capacity: float
    electrical_capacity: float
    shaft_generator_state: str


class MachineryMode:                              # ship_engine.py:23-44
    def __init__(self, params: MachineryModeParams):
        self.main_engine_capacity = params.main_engine_capacity
        self.electrical_capacity = params.electrical_capacity
        self.shaft_generator_state = params.shaft_generator_state
        self.available_propulsion_power = 0
        self.available_propulsion_power_main_engine = 0
        self.available_propulsion_power_electrical = 0

    def update_available_propulsion_power(self, hotel_load):
        if self.shaft_generator_state == 'MOTOR':      # PTI
            self.available_propulsion_power = self.main_engine_capacity + self.electrical_capacity - hotel_load
            self.available_propulsion_power_main_engine = self.main_engine_capacity
            self.available_propulsion_power_electrical = self.electrical_capacity - hotel_load
        elif self.shaft_generator_state == 'GEN':      # PTO
            self.available_propulsion_power = self.main_engine_capacity - hotel_load
            self.available_propulsion_power_main_engine = self.main_engine_capacity - hotel_load
            self.available_propulsion_power_electrical = 0
        else:                                          # MEC
            self.available_propulsion_power = self.main_engine_capacity
            self.available_propulsion_power_main_engine = self.main_engine_capacity
            self.available_propulsion_power_electrical = 0


class MachineryModes:                             # ship_engine.py:78-85
    def __init__(self, list_of_modes: List[MachineryMode]):
        self.list_of_modes = list_of_modes


class FuelConsumptionCoefficients(NamedTuple):    # ship_engine.py:115-118
    a: float
    b: float
    c: float


class SpecificFuelConsumptionWartila6L26:         # ship_engine.py:88-99 (fuel logging is off the step path)
    def __init__(self):
        self.a, self.b, self.c = 128.9, -168.9, 246.8

    def fuel_consumption_coefficients(self):
        return FuelConsumptionCoefficients(a=self.a, b=self.b, c=self.c)


class SpecificFuelConsumptionBaudouin6M26Dot3:    # ship_engine.py:101-112
    def __init__(self):
        self.a, self.b, self.c = 108.7, -289.9, 324.9

    def fuel_consumption_coefficients(self):
        return FuelConsumptionCoefficients(a=self.a, b=self.b, c=self.c)


class MachinerySystemConfiguration(NamedTuple):   # ship_engine.py:121-138
    hotel_load: float
    machinery_modes: MachineryModes
    machinery_operating_mode: int
    rated_speed_main_engine_rpm: float
    linear_friction_main_engine: float
    linear_friction_hybrid_shaft_generator: float
    gear_ratio_between_main_engine_and_propeller: float
    gear_ratio_between_hybrid_shaft_generator_and_propeller: float
    propeller_inertia: float
    propeller_speed_to_torque_coefficient: float
    propeller_diameter: float
    propeller_speed_to_thrust_force_coefficient: float
    rudder_angle_to_sway_force_coefficient: float
    rudder_angle_to_yaw_force_coefficient: float
    max_rudder_angle_degrees: float
    specific_fuel_consumption_coefficients_me: FuelConsumptionCoefficients
    specific_fuel_consumption_coefficients_dg: FuelConsumptionCoefficients


class RudderConfiguration(NamedTuple):            # ship_engine.py:160-163
    rudder_angle_to_sway_force_coefficient: float
    rudder_angle_to_yaw_force_coefficient: float
    max_rudder_angle_degrees: float


class _Integrator:
    """Holder for the integrator step (EulerInt, utils/utils.py:7-53); integration happens on the GPU."""

    def __init__(self, dt=0.01, sim_time=10):
        self.dt = dt
        self.sim_time = sim_time
        self.time = 0.0


class ShipMachineryModel:
    """Parameter holder with the attribute names of ShipMachineryModel (ship_engine.py:341-401)."""

    def __init__(self, machinery_config: MachinerySystemConfiguration,
                 initial_propeller_shaft_speed_rad_per_sec: float, time_step: float):
        self.machinery_modes = machinery_config.machinery_modes
        self.hotel_load = machinery_config.hotel_load
        for mode in self.machinery_modes.list_of_modes:            # ship_engine.py:229-234
            mode.update_available_propulsion_power(self.hotel_load)
        self.mode = self.machinery_modes.list_of_modes[machinery_config.machinery_operating_mode]
        self.c_rudder_v = machinery_config.rudder_angle_to_sway_force_coefficient
        self.c_rudder_r = machinery_config.rudder_angle_to_yaw_force_coefficient
        self.rudder_ang_max = machinery_config.max_rudder_angle_degrees * np.pi / 180
        self.w_rated_me = machinery_config.rated_speed_main_engine_rpm * np.pi / 30
        self.d_me = machinery_config.linear_friction_main_engine
        self.d_hsg = machinery_config.linear_friction_hybrid_shaft_generator
        self.r_me = machinery_config.gear_ratio_between_main_engine_and_propeller
        self.r_hsg = machinery_config.gear_ratio_between_hybrid_shaft_generator_and_propeller
        self.jp = machinery_config.propeller_inertia
        self.kp = machinery_config.propeller_speed_to_torque_coefficient
        self.dp = machinery_config.propeller_diameter
        self.kt = machinery_config.propeller_speed_to_thrust_force_coefficient
        self.shaft_speed_max = 1.1 * self.w_rated_me * self.r_me
        self.omega = initial_propeller_shaft_speed_rad_per_sec
        self.time_step = time_step
        self.int = _Integrator(dt=time_step)
        self._initial_parameters = {'omega': self.omega}
        self.fuel_coeffs_for_main_engine = machinery_config.specific_fuel_consumption_coefficients_me
        self.fuel_coeffs_for_diesel_gen = machinery_config.specific_fuel_consumption_coefficients_dg

    def bookkeeping_columns(self, load_perc, omega, dt_machinery, new_row):
        """The machinery bookkeeping columns of ShipModelAST.simulation_results (rl_env ship_model.py:911-937) from
        the two logged quantities they are functions of -- the commanded load fraction and the shaft speed of every
        log row -- evaluated with the reference's expressions in NumPy (post-processing of the device log, off the
        step path): MachineryMode.distribute_load (ship_engine.py:46-76), fuel_consumption / spec_fuel_cons
        (:250-295) and main_engine_torque (:416-423).

        load_perc, omega: arrays over the log rows; dt_machinery: the machinery integrator's dt (0.01 after the
        first reset(), ship_engine.py:331-333 -- the fuel totals are integrated with it); new_row: bool array, False
        where the row repeats the previous one (store_last_simulation_data adds no fuel)."""
        m = self.mode
        lp = np.asarray(load_perc, dtype=np.float64)
        w = np.asarray(omega, dtype=np.float64)
        total = lp * m.available_propulsion_power
        zeros = np.zeros_like(lp)
        if m.shaft_generator_state == 'MOTOR':
            me = np.minimum(total, m.main_engine_capacity)
            el = total + self.hotel_load - me
            pct_el = el / m.electrical_capacity
            pct_me = zeros if m.main_engine_capacity == 0 else me / m.main_engine_capacity
        elif m.shaft_generator_state == 'GEN':
            el = zeros + min(self.hotel_load, m.electrical_capacity)
            me = total + self.hotel_load - el
            pct_me = me / m.main_engine_capacity
            pct_el = zeros if m.electrical_capacity == 0 else el / m.electrical_capacity
        else:
            me = total
            el = zeros + self.hotel_load
            pct_me = me / m.main_engine_capacity
            pct_el = el / m.electrical_capacity

        def spec(pct, c):
            return (c.a * pct ** 2 + c.b * pct + c.c) / 3.6e9
        rate_me = np.where(me == 0, 0.0, me * spec(pct_me, self.fuel_coeffs_for_main_engine))
        rate_el = np.where(pct_el == 0, 0.0, el * spec(pct_el, self.fuel_coeffs_for_diesel_gen))
        gate = np.asarray(new_row, dtype=bool)
        cons_me = np.cumsum(np.where(gate, rate_me * dt_machinery, 0.0))
        cons_el = np.cumsum(np.where(gate, rate_el * dt_machinery, 0.0))
        cons = np.cumsum(np.where(gate, (rate_me + rate_el) * dt_machinery, 0.0))
        p_me = m.available_propulsion_power_main_engine
        torque = np.minimum(lp * p_me / (w + 0.1), p_me / 5 * np.pi / 30)
        return {
            'commanded load fraction me [-]': pct_me, 'commanded load fraction hsg [-]': pct_el,
            'power me [kw]': me / 1000, 'available power me [kw]': zeros + m.main_engine_capacity / 1000,
            'power electrical [kw]': el / 1000, 'available power electrical [kw]': zeros + m.electrical_capacity / 1000,
            'power [kw]': (el + me) / 1000, 'propulsion power [kw]': (lp * m.available_propulsion_power) / 1000,
            'fuel rate me [kg/s]': rate_me, 'fuel rate hsg [kg/s]': rate_el, 'fuel rate [kg/s]': rate_me + rate_el,
            'fuel consumption me [kg]': cons_me, 'fuel consumption hsg [kg]': cons_el, 'fuel consumption [kg]': cons,
            'motor torque [Nm]': torque,
        }


class SimplifiedPropulsionMachinerySystemConfiguration(NamedTuple):   # ship_engine.py:148-157
    hotel_load: float
    machinery_modes: MachineryModes
    machinery_operating_mode: int
    specific_fuel_consumption_coefficients_me: FuelConsumptionCoefficients
    specific_fuel_consumption_coefficients_dg: FuelConsumptionCoefficients
    thrust_force_dynamic_time_constant: float
    rudder_angle_to_sway_force_coefficient: float
    rudder_angle_to_yaw_force_coefficient: float
    max_rudder_angle_degrees: float


class SimplifiedMachineryModel:
    """Parameter holder with the attribute names of SimplifiedMachineryModel (ship_engine.py:484-519): first-order
    thrust-force dynamics  dT/dt = (-k_thrust * T + load * available_power) / time_constant."""

    def __init__(self, machinery_config: SimplifiedPropulsionMachinerySystemConfiguration, time_step: float,
                 initial_thrust_force: float):
        self.machinery_modes = machinery_config.machinery_modes
        self.hotel_load = machinery_config.hotel_load
        for mode in self.machinery_modes.list_of_modes:
            mode.update_available_propulsion_power(self.hotel_load)
        self.mode = self.machinery_modes.list_of_modes[machinery_config.machinery_operating_mode]
        self.c_rudder_v = machinery_config.rudder_angle_to_sway_force_coefficient
        self.c_rudder_r = machinery_config.rudder_angle_to_yaw_force_coefficient
        self.rudder_ang_max = machinery_config.max_rudder_angle_degrees * np.pi / 180
        self.thrust = initial_thrust_force
        self.d_thrust = 0
        self.k_thrust = 2160 / 790
        self.thrust_time_constant = machinery_config.thrust_force_dynamic_time_constant
        self.time_step = time_step
        self.int = _Integrator(dt=time_step)
