"""Initial-condition helpers with the reference's signatures (NumPy on the host).

construct_primitive_state: initial_condition_generation/construct_primitive_state.py:18-116.
get_helper_data: data_classes/simulation_helper_data.py:49-160 (Cartesian part: cell centres
and distances to the box centre; the production-mode empty tuple is kept).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import option_classes as oc


class HelperData(NamedTuple):
    geometric_centers: object = None
    volumetric_centers: object = None
    r: object = None
    r_hat_alpha: object = None
    cell_volumes: object = None
    inner_cell_boundaries: object = None
    outer_cell_boundaries: object = None


def get_helper_data(config, sharding=None, padded=False, production=False, dtype=np.float64):
    if config.geometry != oc.CARTESIAN:
        raise NotImplementedError("only CARTESIAN geometry is built")
    ngc = config.num_ghost_cells if padded else 0
    dx = config.box_size / config.num_cells
    needed = config.wind_config.stellar_wind or config.cooling_config.cooling or config.return_snapshots
    if production and not needed:
        return HelperData()
    n = config.num_cells + 2 * ngc
    if config.dimensionality > 1:
        x = np.linspace(dx / 2 - ngc * dx, config.box_size + dx / 2 + ngc * dx, n, endpoint=False).astype(dtype)
        grids = np.array(np.meshgrid(*([x] * config.dimensionality)))
        centers = np.moveaxis(grids, 0, -1)
        r = np.linalg.norm(centers - config.box_size / 2, axis=-1)
        return HelperData(geometric_centers=centers, volumetric_centers=centers, r=r)
    r = np.linspace(dx / 2 - ngc * dx, config.box_size - dx / 2 + ngc * dx, n).astype(dtype)
    ones = dx * np.ones_like(r)
    return HelperData(geometric_centers=r, volumetric_centers=r, r_hat_alpha=ones, cell_volumes=ones,
                      inner_cell_boundaries=r - dx / 2, outer_cell_boundaries=r + dx / 2)


def construct_primitive_state(config, registered_variables, density, velocity_x=None, velocity_y=None,
                              velocity_z=None, magnetic_field_x=None, magnetic_field_y=None,
                              magnetic_field_z=None, interface_magnetic_field_x=None,
                              interface_magnetic_field_y=None, interface_magnetic_field_z=None,
                              gas_pressure=None, cosmic_ray_pressure=None, sharding=None, dtype=None):
    """Stack fields into ``[var, x(, y(, z))]``.  dtype defaults to numerical_precision's
    (the reference uses the JAX x64 flag, Q13; pass dtype explicitly to override)."""
    rv = registered_variables
    if dtype is None:
        dtype = np.float64 if config.numerical_precision == oc.DOUBLE_PRECISION else np.float32
    density = np.asarray(density)
    state = np.zeros((rv.num_vars,) + density.shape, dtype=dtype)
    state[rv.density_index] = density
    if config.dimensionality == 1:
        state[rv.velocity_index] = velocity_x
    else:
        state[rv.velocity_index.x] = velocity_x
        state[rv.velocity_index.y] = velocity_y
        if config.dimensionality == 3:
            state[rv.velocity_index.z] = velocity_z
    if config.mhd and config.dimensionality >= 2:
        for idx, f in zip(rv.magnetic_index, (magnetic_field_x, magnetic_field_y, magnetic_field_z)):
            if f is not None:
                state[idx] = f
    state[rv.pressure_index] = gas_pressure
    if rv.cosmic_ray_n_active:
        raise NotImplementedError("cosmic rays are outside the finite-volume hot path")
    return state
