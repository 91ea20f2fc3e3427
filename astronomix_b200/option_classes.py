"""Static and dynamic option carriers of the finite-volume (M)HD step.

Dependency-free mirror (no jax) of the reference's option tuples: same names, field
order, defaults and integer option codes, so reference scripts port by changing the
import.

Reference: astronomix/option_classes/simulation_config.py:20-93 (option codes),
:120-158 (snapshot/boundary settings), :160-343 (SimulationConfig), :345-470
(finalize_config); astronomix/option_classes/simulation_params.py:11-66; the module
option tuples under astronomix/_physics_modules/*/…_options.py (flags only: the
modules themselves are outside the hot path).
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional, Union

# --- option codes (simulation_config.py:20-93) --------------------------------
FINITE_VOLUME, FINITE_DIFFERENCE = 0, 1
FORWARDS, BACKWARDS = 0, 1
MINMOD, OSHER, DOUBLE_MINMOD, SUPERBEE, VAN_ALBADA, VAN_ALBADA_PP = 0, 1, 2, 3, 4, 5
UNSPLIT, SPLIT = 0, 1
HLL, HLLC, HLLC_LM, LAX_FRIEDRICHS, HYBRID_HLLC, AM_HLLC = 0, 1, 2, 3, 4, 5
RK2_SSP, MUSCL, RK4_SSP = 0, 1, 2
OPEN_BOUNDARY, REFLECTIVE_BOUNDARY, PERIODIC_BOUNDARY, MHD_JET_BOUNDARY = 0, 1, 2, 3
GAS_STATE, VELOCITY_ONLY, MAGNETIC_FIELD_ONLY = 0, 1, 2
CARTESIAN, CYLINDRICAL, SPHERICAL = 0, 1, 2
VARAXIS, XAXIS, YAXIS, ZAXIS = 0, 1, 2, 3
GHOST_CELLS, PERIODIC_ROLL = 0, 1
SIMPLE_SOURCE_TERM, DONOR_ACCOUNTING, RIEMANN_SPLIT, RIEMANN_SPLIT_UNSTABLE, HALF_SPLIT = 0, 1, 2, 3, 4
IMPLICIT_MIDPOINT, IMPLICIT_EULER = 0, 1
SINGLE_PRECISION, DOUBLE_PRECISION = 0, 1

# stellar-wind injection schemes (stellar_wind_options.py)
MEO, EI, MEI = 0, 1, 2


# --- flags of the physics modules that sit outside the step -------------------
class WindConfig(NamedTuple):
    stellar_wind: bool = False
    num_injection_cells: int = 10
    wind_injection_scheme: int = EI
    trace_wind_density: bool = False


class WindParams(NamedTuple):
    wind_mass_loss_rate: float = 0.0
    wind_final_velocity: float = 0.0
    pressure_floor: float = 100000.0


class CosmicRayConfig(NamedTuple):
    cosmic_rays: bool = False
    diffusive_shock_acceleration: bool = False


class CosmicRayParams(NamedTuple):
    diffusive_shock_acceleration_start_time: float = 0.0
    diffusive_shock_acceleration_efficiency: float = 0.1


class CoolingConfig(NamedTuple):
    cooling: bool = False
    cooling_method: int = 2
    cooling_curve_config: Any = None


class CoolingParams(NamedTuple):
    hydrogen_mass_fraction: float = 0.76
    metal_mass_fraction: float = 0.02
    floor_temperature: float = 1e4
    cooling_curve_params: Any = None


class TurbulentForcingConfig(NamedTuple):
    turbulent_forcing: bool = False


class TurbulentForcingParams(NamedTuple):
    energy_injection_rate: float = 2.0


class NeuralNetForceConfig(NamedTuple):
    neural_net_force: bool = False
    network_static: Any = None


class NeuralNetForceParams(NamedTuple):
    network_params: Any = None


class CNNMHDconfig(NamedTuple):
    cnn_mhd_corrector: bool = False
    network_static: Any = None


class SnapshotSettings(NamedTuple):
    return_states: bool = True
    return_final_state: bool = False
    return_total_mass: bool = False
    return_total_energy: bool = False
    return_internal_energy: bool = False
    return_kinetic_energy: bool = False
    return_gravitational_energy: bool = False
    return_radial_momentum: bool = False
    return_magnetic_divergence: bool = False


class BoundarySettings1D(NamedTuple):
    left_boundary: int = OPEN_BOUNDARY
    right_boundary: int = OPEN_BOUNDARY


class BoundarySettings(NamedTuple):
    x: BoundarySettings1D = BoundarySettings1D()
    y: BoundarySettings1D = BoundarySettings1D()
    z: BoundarySettings1D = BoundarySettings1D()


class SimulationConfig(NamedTuple):
    """Static options; field-for-field simulation_config.py:160-343."""

    solver_mode: int = FINITE_VOLUME
    numerical_precision: int = SINGLE_PRECISION
    runtime_debugging: bool = False
    donate_state: bool = False
    memory_analysis: bool = False
    print_elapsed_time: bool = False
    progress_bar: bool = False
    dimensionality: int = 1
    state_struct: bool = False
    geometry: int = CARTESIAN
    mhd: bool = False
    fv_magnetic_integrator: int = IMPLICIT_MIDPOINT
    enforce_positivity: bool = True
    self_gravity: bool = False
    self_gravity_version: int = DONOR_ACCOUNTING
    box_size: float = 1.0
    num_cells: int = 400
    reconstruction_order: int = 1
    limiter: int = MINMOD
    riemann_solver: int = HLL
    split: int = UNSPLIT
    time_integrator: int = RK2_SSP
    # evaluated once at class-definition time in the reference, i.e. always 2 (Q8)
    num_ghost_cells: int = 2
    # likewise a definition-time constant (1.0 / 400) until finalize_config runs
    grid_spacing: float = 1.0 / 400
    boundary_handling: int = GHOST_CELLS
    boundary_settings: Union[None, BoundarySettings1D, BoundarySettings] = None
    fixed_timestep: bool = False
    exact_end_time: bool = True
    source_term_aware_timestep: bool = False
    num_timesteps: int = 1000
    use_max_adaptive_timestep: bool = True
    differentiation_mode: int = FORWARDS
    num_checkpoints: int = 100
    return_snapshots: bool = False
    snapshot_settings: SnapshotSettings = SnapshotSettings()
    activate_snapshot_callback: bool = False
    use_specific_snapshot_timepoints: bool = False
    num_snapshots: int = 10
    first_order_fallback: bool = False
    turbulent_forcing_config: TurbulentForcingConfig = TurbulentForcingConfig()
    wind_config: WindConfig = WindConfig()
    cosmic_ray_config: CosmicRayConfig = CosmicRayConfig()
    cooling_config: CoolingConfig = CoolingConfig()
    neural_net_force_config: NeuralNetForceConfig = NeuralNetForceConfig()
    cnn_mhd_corrector_config: CNNMHDconfig = CNNMHDconfig()


class SimulationParams(NamedTuple):
    """Dynamic (traced, in the reference) parameters; simulation_params.py:11-66."""

    C_cfl: float = 0.4
    gravitational_constant: float = 1.0
    gamma: float = 5 / 3
    minimum_density: float = 1e-14
    minimum_pressure: float = 1e-14
    dt_max: float = 0.001
    t_end: float = 0.2
    snapshot_timepoints: Any = (0.0,)
    turbulent_forcing_params: TurbulentForcingParams = TurbulentForcingParams()
    wind_params: WindParams = WindParams()
    cosmic_ray_params: CosmicRayParams = CosmicRayParams()
    cooling_params: CoolingParams = CoolingParams()
    neural_net_force_params: NeuralNetForceParams = NeuralNetForceParams()
    cnn_mhd_corrector_params: CNNMHDconfig = CNNMHDconfig()


_PERIODIC_1D = BoundarySettings1D(PERIODIC_BOUNDARY, PERIODIC_BOUNDARY)
PERIODIC_BOUNDARIES = BoundarySettings(_PERIODIC_1D, _PERIODIC_1D, _PERIODIC_1D)


def finalize_config(config: SimulationConfig, state_shape) -> SimulationConfig:
    """Fix num_cells / grid_spacing / default boundaries from the state shape.

    Same decisions, in the same order, as simulation_config.py:345-470 for every
    branch the finite-volume path can reach (the finite-difference and stellar-wind
    branches only raise / flip flags that this package rejects later).
    """
    num_cells = int(state_shape[-1])
    config = config._replace(num_cells=num_cells)
    if config.dimensionality == 2 and len(set(state_shape[-2:])) != 1:
        raise ValueError("The number of cells in x and y must be equal.")
    if config.dimensionality == 3 and len(set(state_shape[-3:])) != 1:
        raise ValueError("The number of cells in x, y and z must be equal.")
    config = config._replace(grid_spacing=config.box_size / config.num_cells)

    if config.geometry == SPHERICAL:
        # the reference forces HLL + SPLIT + MUSCL + (minmod) here (:365-385)
        config = config._replace(riemann_solver=HLL, split=SPLIT, time_integrator=MUSCL)
        if config.limiter in (VAN_ALBADA, VAN_ALBADA_PP):
            config = config._replace(limiter=MINMOD)
    if config.self_gravity and config.limiter != MINMOD:
        config = config._replace(limiter=MINMOD)
    if config.solver_mode == FINITE_DIFFERENCE:
        raise NotImplementedError(
            "astronomix_b200 covers the finite-volume step only (SURVEY.md §8); "
            "the finite-difference WENO/CT scheme is out of scope."
        )
    if config.boundary_settings is None:
        if config.geometry == CARTESIAN:
            if config.dimensionality == 1:
                config = config._replace(boundary_settings=BoundarySettings1D())
            else:
                config = config._replace(boundary_settings=BoundarySettings())
        elif config.geometry == SPHERICAL and config.dimensionality == 1:
            config = config._replace(
                boundary_settings=BoundarySettings1D(REFLECTIVE_BOUNDARY, OPEN_BOUNDARY)
            )
    if config.wind_config.stellar_wind:
        config = config._replace(source_term_aware_timestep=True)
    return config


def boundary_codes(config: SimulationConfig):
    """[(left, right)] per spatial axis, from either boundary-settings flavour."""
    bs = config.boundary_settings
    if bs is None:
        raise ValueError("boundary_settings unset: call finalize_config first")
    if config.dimensionality == 1:
        if isinstance(bs, BoundarySettings):
            bs = bs.x
        return [(bs.left_boundary, bs.right_boundary)]
    axes = [bs.x, bs.y, bs.z][: config.dimensionality]
    return [(a.left_boundary, a.right_boundary) for a in axes]
