"""astronomix_b200 -- B200-native finite-volume (M)HD time step behind astronomix's API.

Public names follow astronomix/__init__.py:1-38 for everything on the hot path.
"""
from .option_classes import (  # noqa: F401
    SimulationConfig, SimulationParams, finalize_config, BoundarySettings, BoundarySettings1D,
    FORWARDS, BACKWARDS, MINMOD, OSHER, DOUBLE_MINMOD, SUPERBEE, VAN_ALBADA, VAN_ALBADA_PP,
    HLL, HLLC, HLLC_LM, LAX_FRIEDRICHS, HYBRID_HLLC, AM_HLLC,
    OPEN_BOUNDARY, REFLECTIVE_BOUNDARY, PERIODIC_BOUNDARY, CARTESIAN, CYLINDRICAL, SPHERICAL,
    UNSPLIT, SPLIT, RK2_SSP, MUSCL, GHOST_CELLS, PERIODIC_ROLL, IMPLICIT_MIDPOINT, IMPLICIT_EULER,
    SINGLE_PRECISION, DOUBLE_PRECISION, FINITE_VOLUME, FINITE_DIFFERENCE, WindParams, SnapshotSettings,
)
from .registered_variables import RegisteredVariables, get_registered_variables  # noqa: F401
from .helpers import get_helper_data, construct_primitive_state, HelperData  # noqa: F401


from .stepping import time_integration, StateStruct, SlabSharding  # noqa: F401,E402
from .snapshots import SnapshotData  # noqa: F401,E402
