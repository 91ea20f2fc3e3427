"""C-ABI checks that need no GPU: the library loads, exports every symbol include/amx.h
declares, validates descriptors, and refuses to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from astronomix_b200 import _lib, problems as pr
from astronomix_b200 import option_classes as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "amx.h")).read()
    return sorted(set(re.findall(r"AMX_API\s+[\w\s\*]+?\b(amx_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.PROTOTYPES) == syms
    assert lib.amx_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    # 4*(1+1+3+1+1+1+1+1+1+1+6+1+1+1+1+1+1+8) = 128 bytes of int32, then 5 doubles
    assert C.sizeof(_lib.AmxDesc) == 128 + 40
    assert _lib.AmxDesc.dx.offset == 128
    assert C.sizeof(_lib.AmxStats) == 64


def _desc(**kw):
    dim = kw.pop("dim", 3)
    config, params, rv = pr.make(dim, 16, **kw)
    return _lib.make_desc(config, params, (16,) * dim, _lib.AMX_F32)


def test_workspace_query_and_validation():
    lib = _lib.load()
    n = C.c_size_t()
    d = _desc(mhd=True)
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 0
    cells = 16 ** 3 * 4
    assert n.value >= (2 * 8 + 2 * 5 + 2 * 6) * cells
    d.num_ghost = 3
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 2
    assert b"num_ghost" in lib.amx_last_error()
    d = _desc()
    d.riemann = oc.HYBRID_HLLC  # unsplit, split sweeps and the unsplit VAN_ALBADA_PP reconstruction
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 0
    d.split, d.time_integrator = oc.SPLIT, oc.MUSCL
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 0
    d.split, d.time_integrator, d.limiter = oc.UNSPLIT, oc.RK2_SSP, oc.VAN_ALBADA_PP
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 0
    d.riemann = 6
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 1
    d = _desc()
    d.abi_version = 99
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 1
    d = _desc(dim=2)
    d.bc[1][0] = _lib.AMX_BC_HALO
    assert lib.amx_workspace_bytes(C.byref(d), C.byref(n)) == 1


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_lib.AmxError):
        _lib.require_device()
    import astronomix_b200 as ax
    config, params, rv = pr.make(1, 101, boundaries="open")
    with pytest.raises(_lib.AmxError):
        ax.time_integration(pr.sod(101), config, params, rv)
    lib = _lib.load()
    d = _lib.make_desc(config, params, (101,), _lib.AMX_F32)
    buf = np.zeros(16, np.float32)
    p = buf.ctypes.data_as(C.c_void_p)
    assert lib.amx_cfl(C.byref(d), p, p, p, None) == 3


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "astronomix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "fv_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_config_mirror_defaults():
    """Defaults that shape the path."""
    c = oc.SimulationConfig()
    assert (c.split, c.time_integrator, c.limiter, c.riemann_solver) == (oc.UNSPLIT, oc.RK2_SSP, oc.MINMOD, oc.HLL)
    assert c.boundary_handling == oc.GHOST_CELLS and c.num_ghost_cells == 2
    p = oc.SimulationParams()
    assert (p.C_cfl, p.dt_max, p.t_end) == (0.4, 0.001, 0.2) and abs(p.gamma - 5 / 3) < 1e-15
    import astronomix_b200 as ax
    for dim, mhd, nv, pi in [(1, False, 3, 2), (2, False, 4, 3), (2, True, 7, 3), (3, False, 5, 4), (3, True, 8, 4)]:
        rv = ax.get_registered_variables(oc.SimulationConfig(dimensionality=dim, mhd=mhd))
        assert (rv.num_vars, rv.pressure_index) == (nv, pi)
    with pytest.raises(ValueError):
        ax.finalize_config(oc.SimulationConfig(dimensionality=2), (4, 8, 16))


def test_ghost_shell_workspace_and_jet_validation():
    """Host logic of the ghost shell: only unsplit MHD with a reflective side, and any
    run with the MHD jet inflow, reserve the padded gas buffer; the jet code is accepted only where the
    reference's boundary handler has a branch for it."""
    lib = _lib.load()

    def ws(**kw):
        n = C.c_size_t()
        d = _desc(**kw)
        rc = lib.amx_workspace_bytes(C.byref(d), C.byref(n))
        return rc, n.value

    shell3 = 256 + 5 * 20 ** 3 * 4  # stamp + 5 gas rows of the padded 16^3 grid, fp32
    assert ws(mhd=True, boundaries="reflective")[1] - ws(mhd=True, boundaries="open")[1] == shell3
    assert ws(mhd=True, boundaries="open")[1] == ws(mhd=True)[1]
    assert ws(boundaries="reflective")[1] == ws()[1]                                    # hydro: ghost cells stay images
    split = dict(split=oc.SPLIT, time_integrator=oc.MUSCL)
    assert ws(mhd=True, boundaries="reflective", **split)[1] == ws(mhd=True, **split)[1]  # sweeps end with a fill
    b = oc.BoundarySettings1D
    jet = oc.BoundarySettings(b(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY), b(oc.MHD_JET_BOUNDARY, oc.OPEN_BOUNDARY),
                              b(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY))
    for kw in (dict(mhd=True), dict(), dict(mhd=True, **split)):
        rc, n_jet = ws(dim=2, boundaries=jet, **kw)
        assert rc == 0
        assert n_jet - ws(dim=2, boundaries="open", **kw)[1] == 256 + 4 * 20 ** 2 * 4
    assert ws(dim=2, boundaries=jet, riemann_solver=oc.LAX_FRIEDRICHS)[0] == 2          # AMX_EUNSUPPORTED
    assert b"jet" in lib.amx_last_error()
    wrong_side = oc.BoundarySettings(b(oc.MHD_JET_BOUNDARY, oc.OPEN_BOUNDARY), b(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY),
                                     b(oc.OPEN_BOUNDARY, oc.OPEN_BOUNDARY))
    assert ws(dim=2, boundaries=wrong_side)[0] == 2
    assert ws(dim=3, boundaries=jet)[0] == 2
