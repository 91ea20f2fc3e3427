"""astronomix_b200's option carriers against the reference's own classes, recorded as data by
``oracle/gen_reference_api.py`` (``tests/golden/reference_api.json``): field names, field order and
defaults of every NamedTuple, the integer option codes, the variable registry of every
(dimensionality, mhd) pair, what ``finalize_config`` fills in and which inputs it rejects.  A script
written against the reference must find the same names with the same meaning here."""
import json
import os

import numpy as np
import pytest

import astronomix_b200 as ax
from astronomix_b200 import option_classes as oc
from astronomix_b200 import registered_variables as orv

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_api.json")) as f:
    API = json.load(f)


# carriers of physics modules outside the finite-volume path (out of scope): their on/off
# switch is mirrored (turning one on raises NotImplementedError), their inner parameter tables are not
OPAQUE = ("cooling_config.cooling_curve_config", "cooling_params.cooling_curve_params")


def same(mine, ref, path=""):
    """Recursive comparison of one of our objects with the recorded reference value; returns the list
    of differences (paths)."""
    if path.endswith(OPAQUE):
        return []
    if isinstance(ref, dict) and "__fields__" in ref:
        if not hasattr(mine, "_fields"):
            return [f"{path}: reference has a NamedTuple, here {type(mine).__name__}"]
        if list(mine._fields) != ref["__fields__"]:
            return [f"{path}: fields {list(mine._fields)} != {ref['__fields__']}"]
        out = []
        for f_ in ref["__fields__"]:
            out += same(getattr(mine, f_), ref[f_], f"{path}.{f_}")
        return out
    if isinstance(ref, dict) and "__array__" in ref:
        return [] if np.array_equal(np.asarray(mine, dtype=float).ravel(), np.asarray(ref["__array__"], dtype=float).ravel()) \
            else [f"{path}: {mine!r} != array {ref['__array__']}"]
    if isinstance(ref, dict) and "__repr__" in ref:
        return []  # an object without a data representation (a callable, a jax type): not comparable as data
    if isinstance(ref, list):
        if len(ref) != len(mine):
            return [f"{path}: length {len(mine)} != {len(ref)}"]
        out = []
        for k, (a, b) in enumerate(zip(mine, ref)):
            out += same(a, b, f"{path}[{k}]")
        return out
    if isinstance(ref, float) and isinstance(mine, (int, float)):
        return [] if float(mine) == ref else [f"{path}: {mine!r} != {ref!r}"]
    return [] if (mine == ref and type(mine) is type(ref)) or (mine is None and ref is None) \
        else [f"{path}: {mine!r} != {ref!r}"]


def test_option_codes():
    missing = {n: v for n, v in API["constants"].items() if getattr(oc, n, None) != v}
    assert not missing, f"option constants that differ from the reference's: {missing}"


@pytest.mark.parametrize("name,cls", [("SimulationConfig", oc.SimulationConfig), ("SimulationParams", oc.SimulationParams),
                                      ("BoundarySettings", oc.BoundarySettings),
                                      ("BoundarySettings1D", oc.BoundarySettings1D),
                                      ("SnapshotSettings", oc.SnapshotSettings)])
def test_fields_order_and_defaults(name, cls):
    diffs = same(cls(), API[name], name)
    assert not diffs, "\n".join(diffs)


@pytest.mark.parametrize("key", sorted(API["registered_variables"]))
def test_variable_registry(key):
    dim, kind = int(key[0]), key[3:]
    config = oc.SimulationConfig(dimensionality=dim, mhd=kind == "mhd", progress_bar=False)
    ref = API["registered_variables"][key]
    if "__raises__" in ref:
        with pytest.raises(Exception) as e:
            orv.get_registered_variables(config)
        assert type(e.value).__name__ == ref["__raises__"]
        return
    diffs = same(orv.get_registered_variables(config), ref, key)
    assert not diffs, "\n".join(diffs)


@pytest.mark.parametrize("key", sorted(API["finalize_config"]))
def test_finalize_config(key):
    rec = API["finalize_config"][key]
    config = oc.SimulationConfig(progress_bar=False, **rec["kwargs"])
    if "raises" in rec:
        with pytest.raises(Exception) as e:
            oc.finalize_config(config, tuple(rec["shape"]))
        assert type(e.value).__name__ == rec["raises"]
        return
    diffs = same(oc.finalize_config(config, tuple(rec["shape"])), rec["config"], key)
    assert not diffs, "\n".join(diffs)


def test_public_names():
    """the names a reference script imports from the package root (astronomix/__init__.py:1-38, FV path)"""
    for n in ("time_integration", "SimulationConfig", "SimulationParams", "get_registered_variables",
              "finalize_config", "construct_primitive_state", "get_helper_data", "SnapshotData", "SnapshotSettings"):
        assert hasattr(ax, n), n


@pytest.mark.parametrize("key", sorted(API["construct_and_helpers"]))
def test_construct_primitive_state_and_helper_data(key):
    """construct_primitive_state (initial_condition_generation/construct_primitive_state.py:12-110) and
    get_helper_data (data_classes/simulation_helper_data.py:49-120) against the reference's outputs."""
    rec = API["construct_and_helpers"][key]
    dim, mhd = int(key[0]), key.endswith("mhd")
    config = oc.SimulationConfig(dimensionality=dim, mhd=mhd, progress_bar=False, box_size=2.0, num_cells=4)
    rv = orv.get_registered_variables(config)
    fields = {k: np.asarray(v) for k, v in rec["fields"].items()}
    state = ax.construct_primitive_state(config, rv, dtype=np.float64, **fields)
    assert np.array_equal(state, np.asarray(rec["state"]))
    for tag, padded in (("helper", False), ("helper_padded", True)):
        h = ax.get_helper_data(config, padded=padded)
        for name, ref in rec[tag].items():
            got = getattr(h, name)
            assert got is not None, f"{tag}.{name} missing"
            np.testing.assert_allclose(np.asarray(got), np.asarray(ref), rtol=1e-14, atol=1e-14, err_msg=f"{tag}.{name}")
