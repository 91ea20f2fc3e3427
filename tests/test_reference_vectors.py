"""Golden vectors recorded from the reference's own sources (``oracle/gen_reference_vectors.py``:
the unmodified files of /root/reference/astronomix executed on NumPy through ``oracle/jaxshim``).

CPU leg: the restatement ``oracle/fv_oracle.py`` must reproduce every vector (bit for bit in
almost all cases; the bound below is 1e-13 fp64 / 2e-6 fp32 relative L1 to leave room for
one-ulp differences in how a scalar was rounded).  GPU leg: the CUDA path through the C-ABI
must reproduce them within the north-star tolerance (1e-12 fp64, 1e-5 fp32)."""
import glob
import json
import os

import numpy as np
import pytest

import astronomix_b200 as ax
from astronomix_b200 import option_classes as oc
from astronomix_b200 import problems as pr
from oracle import fv_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
VEC_DIR = os.path.join(HERE, "golden", "reference_vectors")
FILES = sorted(glob.glob(os.path.join(VEC_DIR, "*.npz")))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def load(name):
    z = np.load(os.path.join(VEC_DIR, name + ".npz"))
    return z["initial_state"], z["final_state"], json.loads(str(z["spec"]))


def build(spec):
    """The reference-facing option classes of this package, filled from the recorded spec."""
    b1 = [oc.BoundarySettings1D(l, r) for l, r in spec["boundaries"]]
    bs = b1[0] if spec["dim"] == 1 else oc.BoundarySettings(*(b1 + [b1[-1]] * (3 - len(b1))))
    cfg = dict(spec["cfg"])
    cfg.setdefault("numerical_precision", oc.SINGLE_PRECISION)  # the reference's default
    params = ax.SimulationParams(t_end=spec["t_end"], **spec["params"])
    dtype = np.float64 if spec["dtype"] == "f64" else np.float32
    return pr.make(spec["dim"], spec["n"], mhd=spec["mhd"], box=spec["box"], boundaries=bs, dtype=dtype,
                   params=params, **cfg)


def test_vectors_present():
    with open(os.path.join(VEC_DIR, "INDEX.json")) as f:
        index = json.load(f)
    assert sorted(index["cases"]) == NAMES and len(NAMES) >= 100


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_vector(name):
    s0, ref, spec = load(name)
    config, params, rv = build(spec)
    out = fo.time_integration(s0, fo.Opts.from_config(config, params))
    assert out.dtype == ref.dtype and out.shape == ref.shape
    err = pr.rel_l1(out, ref, spec["dim"], spec["mhd"])
    assert err <= (1e-13 if spec["dtype"] == "f64" else 2e-6), f"{name}: rel-L1 {err:.3e}"


# HYBRID_HLLC / AM_HLLC / VAN_ALBADA_PP run on the generic kernels only in unsplit mode; everything
# recorded here is inside the C-ABI's supported set.
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference_vector(cuda, name):
    from astronomix_b200 import stepping as ti
    s0, ref, spec = load(name)
    config, params, rv = build(spec)
    out = ax.time_integration(s0, config, params, rv)
    out = np.asarray(out)
    assert out.dtype == ref.dtype and out.shape == ref.shape and np.isfinite(out).all()
    err = pr.rel_l1(out, ref, spec["dim"], spec["mhd"])
    assert err <= (1e-12 if spec["dtype"] == "f64" else 1e-5), f"{name}: rel-L1 {err:.3e}"
