"""Variable registry: which row of the ``[var, x, y, z]`` state holds which field.

Mirror of astronomix/variable_registry/registered_variables.py:26-113 (types) and
:116-238 (get_registered_variables). Layouts: 1-D [rho,u,p]; 2-D [rho,vx,vy,p];
2-D MHD [rho,vx,vy,p,Bx,By,Bz] (no vz); 3-D [rho,vx,vy,vz,p]; 3-D MHD
[rho,vx,vy,vz,p,Bx,By,Bz]. The conserved state reuses the slots (momentum in the
velocity rows, total energy in the pressure row).
"""
from __future__ import annotations

from typing import NamedTuple, Union

from .option_classes import FINITE_DIFFERENCE, XAXIS, YAXIS, ZAXIS, SimulationConfig


class StaticIntVector(NamedTuple):
    x: int
    y: int
    z: int


class AxisInfo(NamedTuple):
    axis_in_array: int
    velocity_index: int
    magnetic_index: int


class RegisteredVariables(NamedTuple):
    axes_info: tuple = ()
    num_vars: int = 3
    density_index: int = 0
    velocity_index: Union[int, StaticIntVector] = 1
    momentum_index: Union[int, StaticIntVector] = 1
    magnetic_index: Union[int, StaticIntVector] = -1
    interface_magnetic_field_index: Union[int, StaticIntVector] = -1
    pressure_index: int = 2
    energy_index: int = 2
    wind_density_index: int = -1
    wind_density_active: bool = False
    cosmic_ray_n_index: int = -1
    cosmic_ray_n_active: bool = False


def get_registered_variables(config: SimulationConfig) -> RegisteredVariables:
    rv = RegisteredVariables()
    dim = config.dimensionality
    if dim == 2:
        rv = rv._replace(num_vars=4, velocity_index=StaticIntVector(1, 2, -1), pressure_index=3)
        if config.mhd:
            rv = rv._replace(magnetic_index=StaticIntVector(4, 5, 6), num_vars=7)
    elif dim == 3:
        rv = rv._replace(num_vars=5, velocity_index=StaticIntVector(1, 2, 3), pressure_index=4)
        if config.mhd:
            rv = rv._replace(magnetic_index=StaticIntVector(5, 6, 7), num_vars=8)
    if config.mhd and config.solver_mode == FINITE_DIFFERENCE:
        n = rv.num_vars
        rv = rv._replace(interface_magnetic_field_index=StaticIntVector(n, n + 1, n + 2), num_vars=n + 3)
    if config.wind_config.trace_wind_density:
        rv = rv._replace(wind_density_index=rv.num_vars, num_vars=rv.num_vars + 1, wind_density_active=True)
    if config.cosmic_ray_config.cosmic_rays:
        rv = rv._replace(cosmic_ray_n_index=rv.num_vars, num_vars=rv.num_vars + 1, cosmic_ray_n_active=True)
    if dim == 3 and config.mhd:
        v, b = rv.velocity_index, rv.magnetic_index
        rv = rv._replace(
            axes_info=(AxisInfo(XAXIS, v.x, b.x), AxisInfo(YAXIS, v.y, b.y), AxisInfo(ZAXIS, v.z, b.z))
        )
    return rv._replace(momentum_index=rv.velocity_index, energy_index=rv.pressure_index)
