"""LOS guidance parameter/route holder.  Mirrors */sub_systems/LOS_guidance.py:15-81 (LosParameters,
NavigationSystem.__init__/load_waypoints); next_wpt / los_guidance run in the CUDA kernel."""
from __future__ import annotations

from typing import NamedTuple

import numpy as np


class LosParameters(NamedTuple):
    radius_of_acceptance: float
    lookahead_distance: float
    integral_gain: float
    integrator_windup_limit: float


class NavigationSystem:
    def __init__(self, route, radius_of_acceptance=600, lookahead_distance=450, integral_gain=0.01,
                 integrator_windup_limit=0.5):
        self.route = route
        self.ra = radius_of_acceptance
        self.r = lookahead_distance
        self.ki = integral_gain
        self.e_ct = 0.0
        self.e_ct_int = 0.0
        self.integrator_limit = integrator_windup_limit
        self.load_waypoints(self.route)

    def load_waypoints(self, route, print_init_msg=False):
        # a str is a route file with one "north east" pair per line (np.loadtxt raises OSError /
        # FileNotFoundError exactly as in the reference); anything else is taken as an array
        if isinstance(route, str):
            self.data = np.loadtxt(route)
        else:
            self.data = route
        self.north = []
        self.east = []
        for i in range(0, (int(np.size(self.data) / 2))):
            self.north.append(self.data[i][0])
            self.east.append(self.data[i][1])
