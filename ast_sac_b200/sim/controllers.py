"""Controller parameter holders.  Mirror */sub_systems/controllers.py of the reference (gains
NamedTuples :16-39, PiController :45-54, PidController :94-104, ThrustFromSpeedSetPoint :156-173,
EngineThrottleFromSpeedSetPoint :197-221, HeadingByReferenceController :272-284,
HeadingByRouteController :316-342, HeadingBySampledRouteController :384-422).  The control laws run
in the CUDA kernel; these objects only carry gains, limits and the route."""
from __future__ import annotations

from typing import NamedTuple

from .LOS_guidance import LosParameters, NavigationSystem  # noqa: F401  (re-exported like the reference)


class ThrottleControllerGains(NamedTuple):
    kp_ship_speed: float
    ki_ship_speed: float
    kp_shaft_speed: float
    ki_shaft_speed: float


class SpeedControllerGains(NamedTuple):
    kp: float
    kd: float
    ki: float


class HeadingControllerGains(NamedTuple):
    kp: float
    kd: float
    ki: float


class PiController:
    def __init__(self, kp: float, ki: float, time_step: float, initial_integral_error=0):
        self.kp = kp
        self.ki = ki
        self.error_i = initial_integral_error
        self.time_step = time_step
        self._initial_state = {'kp': kp, 'ki': ki, 'error_i': initial_integral_error, 'time_step': time_step}


class PidController:
    def __init__(self, kp: float, kd: float, ki: float, time_step: float):
        self.kp = kp
        self.kd = kd
        self.ki = ki
        self.error_i = 0
        self.prev_error = 0
        self.time_step = time_step


class ThrustFromSpeedSetPoint:
    def __init__(self, gains: SpeedControllerGains, max_thrust: float, time_step: float):
        self.ship_speed_controller = PidController(kp=gains.kp, ki=gains.ki, kd=gains.kd, time_step=time_step)
        self.max_thrust = max_thrust


class EngineThrottleFromSpeedSetPoint:
    def __init__(self, gains: ThrottleControllerGains, max_shaft_speed: float, time_step: float,
                 initial_shaft_speed_integral_error: float):
        self.ship_speed_controller = PiController(kp=gains.kp_ship_speed, ki=gains.ki_ship_speed, time_step=time_step)
        self.shaft_speed_controller = PiController(kp=gains.kp_shaft_speed, ki=gains.ki_shaft_speed,
                                                   time_step=time_step,
                                                   initial_integral_error=initial_shaft_speed_integral_error)
        self.max_shaft_speed = max_shaft_speed


class ThrottleFromSpeedSetPointSimplifiedPropulsion:
    """rl_env controllers.py:212-232: PI on the ship speed -> throttle in [0, 1.1]."""

    def __init__(self, kp: float, ki: float, time_step: float):
        self.ship_speed_controller = PiController(kp=kp, ki=ki, time_step=time_step)


class HeadingByReferenceController:
    def __init__(self, gains: HeadingControllerGains, time_step, max_rudder_angle):
        self.gains = gains
        self.time_step = time_step
        self.ship_heading_controller = PidController(kp=gains.kp, kd=gains.kd, ki=gains.ki, time_step=time_step)
        self.max_rudder_angle = max_rudder_angle


class HeadingByRouteController:
    def __init__(self, route_name, heading_controller_gains: HeadingControllerGains,
                 los_parameters: LosParameters, time_step: float, max_rudder_angle: float):
        self.heading_controller = HeadingByReferenceController(
            gains=heading_controller_gains, time_step=time_step, max_rudder_angle=max_rudder_angle)
        self.navigate = NavigationSystem(
            route=route_name,
            radius_of_acceptance=los_parameters.radius_of_acceptance,
            lookahead_distance=los_parameters.lookahead_distance,
            integral_gain=los_parameters.integral_gain,
            integrator_windup_limit=los_parameters.integrator_windup_limit)
        self.next_wpt = 1
        self.prev_wpt = 0
        self.heading_ref = 0
        self.heading_mea = 0


class HeadingBySampledRouteController(HeadingByRouteController):
    def __init__(self, route_name, heading_controller_gains: HeadingControllerGains,
                 los_parameters: LosParameters, time_step: float, max_rudder_angle: float,
                 num_of_samplings: int):
        super().__init__(route_name, heading_controller_gains, los_parameters, time_step, max_rudder_angle)
        self.num_of_samplings = num_of_samplings
