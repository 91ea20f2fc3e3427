"""ctypes binding of libamx.so (include/amx.h) -- the only door to the CUDA step.

There is deliberately no fallback: if the shared library is missing or no CUDA device is
usable, every compute call raises.  Build the library with ``__graft_entry__.build()``
(or ``make -C astronomix_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

from . import option_classes as oc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamx.so")
ABI_VERSION = 3

AMX_F32, AMX_F64 = 0, 1
AMX_ARITH_FAST, AMX_ARITH_STRICT = 0, 1
AMX_BC_HALO = 4
ERRORS = {1: "AMX_EINVAL", 2: "AMX_EUNSUPPORTED", 3: "AMX_ECUDA", 4: "AMX_ENOTCONVERGED"}


class AmxDesc(C.Structure):
    """struct amx_desc (include/amx.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("ndim", C.c_int32),
        ("n", C.c_int32 * 3),
        ("dtype", C.c_int32),
        ("mhd", C.c_int32),
        ("riemann", C.c_int32),
        ("limiter", C.c_int32),
        ("split", C.c_int32),
        ("time_integrator", C.c_int32),
        ("boundary_handling", C.c_int32),
        ("bc", (C.c_int32 * 2) * 3),
        ("num_ghost", C.c_int32),
        ("magnetic_integrator", C.c_int32),
        ("first_order_fallback", C.c_int32),
        ("exact_end_time", C.c_int32),
        ("use_max_adaptive_timestep", C.c_int32),
        ("arithmetic", C.c_int32),
        ("peer_flags", C.c_int32),
        ("slab_flags", C.c_int32),
        ("peer_delta_lo", C.c_int64),
        ("peer_delta_hi", C.c_int64),
        ("reserved", C.c_int32 * 2),
        ("dx", C.c_double),
        ("gamma", C.c_double),
        ("C_cfl", C.c_double),
        ("dt_max", C.c_double),
        ("fp_tol", C.c_double),
    ]


class AmxStats(C.Structure):
    """struct amx_stats (include/amx.h)."""

    _fields_ = [
        ("steps", C.c_int64),
        ("magnetic_iterations", C.c_int64),
        ("time", C.c_double),
        ("last_dt", C.c_double),
        ("kernel_launches", C.c_int64),
        ("pair_launches", C.c_int64),
        ("single_launches", C.c_int64),
        ("final_launches", C.c_int64),
    ]


class AmxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


_P = C.c_void_p
_D = C.POINTER(AmxDesc)
# name -> (restype, argtypes): every symbol include/amx.h declares
PROTOTYPES = {
    "amx_abi_version": (C.c_int, []),
    "amx_last_error": (C.c_char_p, []),
    "amx_device_check": (C.c_int, []),
    "amx_profile_enable": (C.c_int, [C.c_int]),
    "amx_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "amx_workspace_bytes": (C.c_int, [_D, C.POINTER(C.c_size_t)]),
    "amx_pad": (C.c_int, [_D, _P, _P, _P]),
    "amx_unpad": (C.c_int, [_D, _P, _P, _P]),
    "amx_boundary_fill": (C.c_int, [_D, _P, C.c_int, C.c_int, _P]),
    "amx_cfl": (C.c_int, [_D, _P, _P, _P, _P]),
    "amx_step": (C.c_int, [_D, _P, _P, _P, _P, _P, C.POINTER(AmxStats)]),
    "amx_integrate": (C.c_int, [_D, _P, _P, C.c_double, C.c_int64, _P, _P, C.POINTER(AmxStats)]),
    "amx_integrate_from": (C.c_int, [_D, _P, _P, C.c_double, C.c_double, C.c_double, C.c_int64, _P, _P,
                                     C.POINTER(AmxStats)]),
    "amx_alpha": (C.c_int, [_D, _P, _P, _P, _P]),
    "amx_dt_from_alpha": (C.c_int, [_D, _P, _P, C.c_double, _P, _P]),
    "amx_hydro_stage": (C.c_int, [_D, _P, _P, _P, _P, C.c_double, _P, _P, _P, _P]),
    "amx_split_sweep": (C.c_int, [_D, _P, _P, C.c_int, _P, C.c_double, _P, _P, _P, _P]),
    "amx_magnetic_iter": (C.c_int, [_D, _P, _P, _P, _P, _P, _P, _P, _P]),
    "amx_magnetic_final": (C.c_int, [_D, _P, _P, _P, _P, _P, _P]),
    "amx_magnetic_iter_guarded": (C.c_int, [_D, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P]),
    "amx_magnetic_check": (C.c_int, [_D, _P, C.c_int, _P, _P]),
    "amx_magnetic_final_guarded": (C.c_int, [_D, _P, _P, _P, _P, _P, _P, _P, _P]),
    "amx_magnetic_pair_supported": (C.c_int, [_D]),
    "amx_magnetic_pair_guarded": (C.c_int, [_D, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "amx_magnetic_check_pair": (C.c_int, [_D, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "amx_magnetic_fix_guarded": (C.c_int, [_D, _P, _P, _P, _P, _P, C.c_int, _P, _P]),
    "amx_magnetic_state": (C.c_int, [_D, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "amx_threefry_normal": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int64, _P, C.c_int, C.c_int, _P]),
    "amx_threefry_normal_range": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int,
                                            C.c_int, _P]),
    "amx_forcing_spectrum": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                       C.c_double, _P, C.c_int, C.c_int, _P]),
    "amx_forcing_sums": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P, C.c_int, _P]),
    "amx_shell_update": (C.c_int, [_D, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "amx_peer_push_supported": (C.c_int, [_D]),
    "amx_peer_comm_bytes": (C.c_size_t, []),
    "amx_peer_allreduce_max": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, _P, C.c_int, C.c_int, _P]),
}

_lib = None


def load():
    """dlopen libamx.so and type its entry points; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    exp = os.environ.get("AMX_LIB_VARIANT")  # kernel A/B builds: libamx_<variant>.so (csrc/Makefile, EXP=...)
    if exp:
        path = os.path.join(os.path.dirname(LIB_PATH), f"libamx_{exp}.so")
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA library is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root. "
            "astronomix_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.amx_abi_version() != ABI_VERSION:
        raise ImportError("libamx.so ABI version mismatch: rebuild it")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise AmxError(code, load().amx_last_error().decode())


def require_device():
    """Fail loudly when the CUDA path cannot run (no silent fallbacks)."""
    check(load().amx_device_check())


def default_arithmetic():
    v = os.environ.get("AMX_ARITHMETIC", "fast").lower()
    if v not in ("fast", "strict"):
        raise ValueError("AMX_ARITHMETIC must be 'fast' or 'strict'")
    return AMX_ARITH_STRICT if v == "strict" else AMX_ARITH_FAST


def make_desc(config, params, n, dtype_code, arithmetic=None, gamma=None, halo_x=False):
    """amx_desc from the reference's option tuples; n = active cells per axis."""
    d = AmxDesc()
    d.abi_version = ABI_VERSION
    d.ndim = config.dimensionality
    nn = list(n) + [1] * (3 - len(n))
    for a in range(3):
        d.n[a] = int(nn[a])
    d.dtype = dtype_code
    d.mhd = int(bool(config.mhd))
    d.riemann = config.riemann_solver
    d.limiter = config.limiter
    d.split = config.split
    d.time_integrator = config.time_integrator
    d.boundary_handling = config.boundary_handling
    codes = oc.boundary_codes(config)
    for a in range(3):
        l, r = codes[a] if a < len(codes) else (oc.PERIODIC_BOUNDARY, oc.PERIODIC_BOUNDARY)
        if config.boundary_handling == oc.PERIODIC_ROLL:
            l = r = oc.PERIODIC_BOUNDARY  # rolls wrap whatever boundary_settings says
        d.bc[a][0], d.bc[a][1] = l, r
    if halo_x:
        d.bc[0][0] = d.bc[0][1] = AMX_BC_HALO
    d.num_ghost = config.num_ghost_cells
    d.magnetic_integrator = config.fv_magnetic_integrator
    d.first_order_fallback = int(bool(config.first_order_fallback))
    d.exact_end_time = int(bool(config.exact_end_time))
    d.use_max_adaptive_timestep = int(bool(config.use_max_adaptive_timestep))
    d.arithmetic = default_arithmetic() if arithmetic is None else arithmetic
    d.dx = float(config.grid_spacing)
    d.gamma = float(params.gamma if gamma is None else gamma)
    d.C_cfl = float(params.C_cfl)
    d.dt_max = float(params.dt_max)
    d.fp_tol = 1e-5 if config.numerical_precision == oc.SINGLE_PRECISION else 1e-10
    return d
