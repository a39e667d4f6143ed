"""ast_sac_b200 -- B200-native batched ship-in-transit environment (drop-in for the simulator-step
hot path of AndreasKing-Goks/ast-sac).  See DESIGN.md and INTEGRATION.md at the repository root.

(The directory is ``ast_sac_b200`` because a Python package name cannot contain the hyphen of
``ast-sac``.)
"""
from .env import (BatchedShipEnv, MultiShipEnv, MultiShipNonIWEnv, MultiShipRLEnv, ShipAssets,  # noqa: F401
                  events_to_string)
from .sim.controllers import (EngineThrottleFromSpeedSetPoint, HeadingByRouteController,  # noqa: F401
                              HeadingBySampledRouteController, HeadingControllerGains, LosParameters,
                              SpeedControllerGains, ThrottleControllerGains, ThrustFromSpeedSetPoint)
from .sim.obstacle import PolygonObstacle  # noqa: F401
from .sim.ship_engine import (MachineryMode, MachineryModeParams, MachineryModes,  # noqa: F401
                              MachinerySystemConfiguration, RudderConfiguration)
from .sim.ship_model import (EnvironmentConfiguration, ShipConfiguration, ShipModel, ShipModelAST,  # noqa: F401
                             SimpleShipModel, SimulationConfiguration)

__version__ = "0.1.0"
