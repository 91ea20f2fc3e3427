"""Host-side mirror of the reference's public entry point and its two per-step seams.

``time_integration(primitive_state, config, params, registered_variables, ...)`` keeps
the reference's signature (astronomix/time_stepping/time_integration.py:61-68) and
returns the unpadded primitive state; ``_evolve_state_fv`` and ``_cfl_time_step`` keep
the signatures of the functions the loop body calls
(_finite_volume/_state_evolution/evolve_state.py:496-506,
_finite_volume/_timestep_estimation/_timestep_estimator.py:94-103) and operate on the
same (padded, in GHOST_CELLS mode) arrays.  All arithmetic happens in libamx (CUDA);
PyTorch only owns device memory and the stream.  NumPy arrays are staged through pinned
host memory; torch CUDA tensors are used in place.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from . import option_classes as oc
from .registered_variables import RegisteredVariables


class StateStruct(NamedTuple):
    """data_classes/simulation_snapshot_data.py:218 (only the field the FV path touches)."""

    primitive_state: object = None


class IntegrationInfo(NamedTuple):
    steps: int
    time: float
    last_dt: float
    magnetic_iterations: int
    kernel_launches: int
    pair_launches: int = 0    # fixed-point launches that made two evaluations (amx_pair.cu)
    single_launches: int = 0  # fixed-point launches that made one evaluation
    final_launches: int = 0   # magnetic write-back launches that did work


#: statistics of the most recent time_integration call on this process
last_info: Optional[IntegrationInfo] = None

_TORCH_DTYPES = {torch.float32: _lib.AMX_F32, torch.float64: _lib.AMX_F64}


def _reject_unsupported(config, registered_variables=None):
    """Options that leave the hot path of SURVEY.md §8 raise instead of being ignored."""
    if config.solver_mode != oc.FINITE_VOLUME:
        raise NotImplementedError("only solver_mode=FINITE_VOLUME is built (SURVEY.md §8)")
    if config.geometry != oc.CARTESIAN:
        raise NotImplementedError("only CARTESIAN geometry is built (spherical 1-D is out of scope)")
    off = {
        "self_gravity": config.self_gravity,
        "fixed_timestep": config.fixed_timestep,
        "source_term_aware_timestep": config.source_term_aware_timestep,
        "runtime_debugging": config.runtime_debugging,
        "stellar_wind": config.wind_config.stellar_wind,
        "cosmic_rays": config.cosmic_ray_config.cosmic_rays,
        "cooling": config.cooling_config.cooling,
        "neural_net_force": config.neural_net_force_config.neural_net_force,
        "cnn_mhd_corrector": config.cnn_mhd_corrector_config.cnn_mhd_corrector,
    }
    on = [k for k, v in off.items() if v]
    if on:
        raise NotImplementedError(f"options outside the finite-volume hot path are not built: {on}")
    if registered_variables is not None and (
        registered_variables.wind_density_active or registered_variables.cosmic_ray_n_active
    ):
        raise NotImplementedError("extra registered variables (wind density, cosmic rays) are not built")


#: pinned staging buffers for NumPy / CPU-tensor states, by direction ("in" / "out"); allocated once per size
#: (cudaHostAlloc of half a gigabyte costs more than the copy it serves)
_staging = {}


def _pinned(shape, dtype, slot):
    need = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    ent = _staging.get(slot)
    if ent is None or ent[0].numel() < need:
        ent = [torch.empty(max(need, 1), dtype=torch.uint8, pin_memory=True), None]
        _staging[slot] = ent
    if ent[1] is not None:  # the previous transfer out of / into this buffer must have finished
        ent[1].synchronize()
    return ent[0][:need].view(dtype).view(tuple(shape)), ent


def _to_device(state):
    """-> (contiguous CUDA tensor, kind) ; kind remembers how to hand the result back."""
    if isinstance(state, torch.Tensor):
        if state.is_cuda:
            return state.contiguous(), "torch_cuda"
        src, kind = state.contiguous(), "torch_cpu"
    else:
        arr = np.ascontiguousarray(state)
        if arr.dtype not in (np.float32, np.float64):
            raise TypeError(f"state dtype must be float32 or float64, got {arr.dtype}")
        src, kind = torch.from_numpy(arr), "numpy"
    if src.is_pinned():
        return src.cuda(non_blocking=True), kind
    stage, ent = _pinned(src.shape, src.dtype, "in")
    t = torch.empty(src.shape, dtype=src.dtype, device="cuda")
    if src.dim() > 1 and src.shape[0] > 1 and src.numel() > (1 << 22):
        # row by row: the host copy of row r + 1 into the staging buffer overlaps the H2D copy of row r
        for r in range(src.shape[0]):
            stage[r].copy_(src[r])
            t[r].copy_(stage[r], non_blocking=True)
    else:
        stage.copy_(src)
        t.copy_(stage, non_blocking=True)
    ent[1] = torch.cuda.Event()
    ent[1].record()
    return t, kind


def _from_device(t, kind):
    if kind == "torch_cuda":
        return t
    host, ent = _pinned(t.shape, t.dtype, "out")
    out = torch.empty(t.shape, dtype=t.dtype)
    if t.dim() > 1 and t.shape[0] > 1 and t.numel() > (1 << 22):
        # row by row: the D2H copy of row r + 1 overlaps the host copy of row r into the result
        evs = []
        for r in range(t.shape[0]):
            host[r].copy_(t[r], non_blocking=True)
            e = torch.cuda.Event()
            e.record()
            evs.append(e)
        for r in range(t.shape[0]):
            evs[r].synchronize()
            out[r].copy_(host[r])
    else:
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out.copy_(host)
    return out.numpy() if kind == "numpy" else out


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache = {}


def _workspace(desc, device):
    """Device scratch for a descriptor, cached by size (allocation is not part of a step)."""
    lib = _lib.load()
    nbytes = C.c_size_t()
    _lib.check(lib.amx_workspace_bytes(C.byref(desc), C.byref(nbytes)))
    key = (device, nbytes.value)
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    base = ws.data_ptr()
    return ws, C.c_void_p((base + 255) & ~255)


def _check_state(t, nrows, ndim):
    if t.dtype not in _TORCH_DTYPES:
        raise TypeError(f"state dtype must be float32 or float64, got {t.dtype}")
    if t.dim() != ndim + 1 or t.shape[0] != nrows:
        raise ValueError(f"state must be [{nrows}, " + ", ".join(["n"] * ndim) + f"], got {tuple(t.shape)}")


def time_integration(primitive_state, config, params, registered_variables: RegisteredVariables,
                     snapshot_callable=None, sharding=None):
    """Integrate to ``params.t_end``; drop-in for astronomix.time_integration (FV path)."""
    global last_info
    _lib.require_device()
    _reject_unsupported(config, registered_variables)
    if config.activate_snapshot_callback and snapshot_callable is None:
        raise ValueError("activate_snapshot_callback needs a snapshot_callable")
    if sharding is not None:
        return _time_integration_sharded(primitive_state, config, params, registered_variables, sharding)
    struct = config.state_struct
    state = primitive_state.primitive_state if struct else primitive_state
    t, kind = _to_device(state)
    _check_state(t, registered_variables.num_vars, config.dimensionality)
    forced = config.turbulent_forcing_config.turbulent_forcing
    if forced or config.return_snapshots or config.activate_snapshot_callback:
        # snapshots / diagnostics / callback at the reference's snapshot times: segments of
        # the device-resident loop; with turbulent driving (f1) the state is touched between dt and the
        # step, so the loop body runs one step at a time
        if forced:
            from .forcing import integrate_per_step as run
        else:
            from .snapshots import integrate_with_snapshots as run
        final, data, (steps, time, mag_iters, launches) = run(
            t, kind, config, params, registered_variables, snapshot_callable, _from_device)
        last_info = IntegrationInfo(steps, time, 0.0, mag_iters, launches)
        if data is not None:
            return data
        res = _from_device(final, kind)
        return StateStruct(primitive_state=res) if struct else res
    desc = _lib.make_desc(config, params, t.shape[1:], _TORCH_DTYPES[t.dtype])
    ws, ws_ptr = _workspace(desc, t.device)
    out = torch.empty_like(t)
    stats = _lib.AmxStats()
    lib = _lib.load()
    with torch.cuda.device(t.device):
        _lib.check(lib.amx_integrate(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()),
                                     float(params.t_end), -1, ws_ptr, _stream_ptr(), C.byref(stats)))
    last_info = IntegrationInfo(stats.steps, stats.time, stats.last_dt, stats.magnetic_iterations,
                                stats.kernel_launches, stats.pair_launches, stats.single_launches,
                                stats.final_launches)
    res = _from_device(out, kind)
    return StateStruct(primitive_state=res) if struct else res


class SlabSharding(NamedTuple):
    """What ``time_integration(..., sharding=...)`` takes here.  The reference takes a jax ``NamedSharding`` of
    the ``[var, x, y, z]`` state (time_integration.py:61-68, :108-118; the scaling scripts shard axis x:
    arena/arena_tests/scaling/scaling.py:93-98) and lets GSPMD insert the collectives; here the same
    decomposition is explicit: one process per GPU under ``torch.distributed``, x slabs in rank order
    (astronomix_b200.slab).  ``"x"`` and partition tuples such as ``(None, "x", None, None)`` are accepted
    as shorthands for ``SlabSharding()``."""

    group: object = None        # torch.distributed process group (None: the default group)
    local_input: bool = False   # primitive_state is this rank's slab [var, n_x / P, ...], not the global array
    local_output: bool = False  # return this rank's slab instead of the gathered global array


def _as_slab_sharding(sharding):
    if isinstance(sharding, SlabSharding):
        return sharding
    if sharding == "x":
        return SlabSharding()
    if isinstance(sharding, (tuple, list)) and len(sharding) >= 2 and sharding[1] is not None \
            and all(a is None for i, a in enumerate(sharding) if i != 1):
        return SlabSharding()
    raise NotImplementedError(
        f"sharding={sharding!r}: only the x-slab decomposition is built (SlabSharding(), 'x' or (None, 'x', ...))")


def _time_integration_sharded(primitive_state, config, params, registered_variables, sharding):
    """time_integration on an x-slab decomposition: every rank of the group makes the same call."""
    global last_info
    import torch.distributed as dist

    from .slab import SlabDriver
    sh = _as_slab_sharding(sharding)
    if not dist.is_initialized():
        raise RuntimeError("sharding needs torch.distributed (one process per GPU, e.g. under torchrun)")
    if config.return_snapshots or config.activate_snapshot_callback:
        raise NotImplementedError("snapshots / callbacks on a sharded run are not built")
    struct = config.state_struct
    state = primitive_state.primitive_state if struct else primitive_state
    world = dist.get_world_size(sh.group)
    device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(state, torch.Tensor):
        kind = "torch_cuda" if state.is_cuda else "torch_cpu"
        src = state
    else:
        kind, src = "numpy", torch.from_numpy(np.ascontiguousarray(state))
    if src.dtype not in _TORCH_DTYPES:
        raise TypeError(f"state dtype must be float32 or float64, got {src.dtype}")
    shape = tuple(int(v) for v in src.shape[1:])
    global_n = ((shape[0] * world,) + shape[1:]) if sh.local_input else shape
    if src.dim() != config.dimensionality + 1 or src.shape[0] != registered_variables.num_vars:
        raise ValueError(f"state must be [{registered_variables.num_vars}, ...], got {tuple(src.shape)}")
    drv = SlabDriver(config, params, registered_variables, global_n, src.dtype, device, group=sh.group)
    if config.turbulent_forcing_config.turbulent_forcing:
        from .forcing import SlabForcing
        drv.forcing = SlabForcing(config, params, registered_variables, src.dtype, device, drv.rank, drv.world, sh.group)
    if sh.local_input:
        from .slab import NG
        drv.S[0][:, NG:NG + drv.nxl].copy_(src.to(device))
        drv.reset()
        drv.time.zero_()
        drv._init_shell(drv.S[0][: drv.ng])
        drv._alpha_of(drv.S[0][: drv.ng])
    else:
        drv.scatter(src)
    info = drv.run()
    last_info = IntegrationInfo(info.steps, info.time, float(drv.dt.item()), info.magnetic_iterations, info.launches)
    out = drv.local_result().contiguous() if sh.local_output else drv.gather()
    res = out if kind == "torch_cuda" else (out.cpu().numpy() if kind == "numpy" else out.cpu())
    return StateStruct(primitive_state=res) if struct else res


def integrate_steps(primitive_state, config, params, registered_variables, num_steps, arithmetic=None):
    """time_integration truncated after ``num_steps`` steps (parity tests fix the step count)."""
    global last_info
    _lib.require_device()
    _reject_unsupported(config, registered_variables)
    t, kind = _to_device(primitive_state)
    _check_state(t, registered_variables.num_vars, config.dimensionality)
    desc = _lib.make_desc(config, params, t.shape[1:], _TORCH_DTYPES[t.dtype], arithmetic=arithmetic)
    ws, ws_ptr = _workspace(desc, t.device)
    out = torch.empty_like(t)
    stats = _lib.AmxStats()
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_integrate(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()),
                                             float(params.t_end), int(num_steps), ws_ptr, _stream_ptr(),
                                             C.byref(stats)))
    last_info = IntegrationInfo(stats.steps, stats.time, stats.last_dt, stats.magnetic_iterations,
                                stats.kernel_launches, stats.pair_launches, stats.single_launches,
                                stats.final_launches)
    return _from_device(out, kind)


def _active_shape(config, seam_shape):
    g = config.num_ghost_cells if config.boundary_handling == oc.GHOST_CELLS else 0
    return tuple(int(s) - 2 * g for s in seam_shape)


def _cfl_time_step(primitive_state, grid_spacing, dt_max, gamma, config, registered_variables, C_CFL=0.8):
    """dt from the (padded) state; mirror of _timestep_estimator.py:94-218. Returns a 0-d tensor."""
    _lib.require_device()
    t, kind = _to_device(primitive_state)
    params = oc.SimulationParams(C_cfl=float(C_CFL), gamma=float(gamma), dt_max=float(dt_max))
    cfg = config._replace(grid_spacing=float(grid_spacing))
    desc = _lib.make_desc(cfg, params, _active_shape(cfg, t.shape[1:]), _TORCH_DTYPES[t.dtype])
    # (the MHD loop calls the estimator with the full state; only the gas rows are read)
    ws, ws_ptr = _workspace(desc, t.device)
    dt = torch.empty((), dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_cfl(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(dt.data_ptr()),
                                       ws_ptr, _stream_ptr()))
    return dt if kind == "torch_cuda" else dt.cpu().numpy()[()]


def _evolve_state_fv(primitive_state, dt, gamma, gravitational_constant, config, params, helper_data,
                     registered_variables):
    """One step on the (padded) state; mirror of evolve_state.py:496-600."""
    _lib.require_device()
    _reject_unsupported(config, registered_variables)
    t, kind = _to_device(primitive_state)
    _check_state(t, registered_variables.num_vars, config.dimensionality)
    desc = _lib.make_desc(config, params, _active_shape(config, t.shape[1:]), _TORCH_DTYPES[t.dtype], gamma=gamma)
    ws, ws_ptr = _workspace(desc, t.device)
    if isinstance(dt, torch.Tensor) and dt.is_cuda:
        dt_dev = dt.to(t.dtype).reshape(())
    else:
        dt_dev = torch.tensor(float(dt), dtype=t.dtype).to(t.device)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_step(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(dt_dev.data_ptr()), ws_ptr, _stream_ptr(), None))
    return _from_device(out, kind)


def _seam_desc(config, padded_shape, dtype, all_open=False):
    """Descriptor of the active grid behind a padded (seam) array; all_open: every side an open copy."""
    n = _active_shape(config, padded_shape)
    desc = _lib.make_desc(config, oc.SimulationParams(), n, _TORCH_DTYPES[dtype])
    if all_open:
        for a in range(config.dimensionality):
            desc.bc[a][0] = desc.bc[a][1] = oc.OPEN_BOUNDARY
    return desc


def _pad(state, config):
    """Edge-pad by two cells per side and axis; mirror of time_stepping/_utils.py:65-75 (Q8).  In
    PERIODIC_ROLL mode the reference never pads and this returns a copy."""
    _lib.require_device()
    t, kind = _to_device(state)
    ghost = config.boundary_handling == oc.GHOST_CELLS
    g = config.num_ghost_cells if ghost else 0
    out = torch.empty((t.shape[0],) + tuple(int(s) + 2 * g for s in t.shape[1:]), dtype=t.dtype, device=t.device)
    # an open side copies the edge cell into its ghost cells, which is np.pad(mode="edge") row by row
    desc = _seam_desc(config, out.shape[1:], t.dtype, all_open=ghost)
    if t.shape[0] != (config.dimensionality + 2 + (3 if config.mhd else 0)):
        raise ValueError("_pad expects the full state (gas rows, then the three magnetic rows for MHD)")
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_pad(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), _stream_ptr()))
    return _from_device(out, kind)


def _unpad(state, config):
    """Slice the ghost cells off; mirror of time_stepping/_utils.py:19-61."""
    _lib.require_device()
    t, kind = _to_device(state)
    desc = _seam_desc(config, t.shape[1:], t.dtype)
    out = torch.empty((t.shape[0],) + _active_shape(config, t.shape[1:]), dtype=t.dtype, device=t.device)
    if t.shape[0] != (config.dimensionality + 2 + (3 if config.mhd else 0)):
        raise ValueError("_unpad expects the full state (gas rows, then the three magnetic rows for MHD)")
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_unpad(C.byref(desc), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), _stream_ptr()))
    return _from_device(out, kind)


def _boundary_handler(primitive_state, config, type_handled=0):
    """Ghost-cell fill of a padded array; mirror of _geometry/boundaries.py:74-441.  type_handled: 0 gas
    state, 1 velocity-only, 2 magnetic-field-only (3-row arrays; the flip rule of SURVEY.md Q12 and the
    MHD jet constants depend on it).  Returns the filled array (the input is not modified unless it is a
    CUDA tensor, which is filled in place like the reference's donated buffer)."""
    _lib.require_device()
    t, kind = _to_device(primitive_state)
    desc = _seam_desc(config, t.shape[1:], t.dtype)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().amx_boundary_fill(C.byref(desc), C.c_void_p(t.data_ptr()), int(t.shape[0]),
                                                 int(type_handled), _stream_ptr()))
    return _from_device(t, kind)
