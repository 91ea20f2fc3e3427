"""Run the UNMODIFIED astronomix sources on NumPy -- TEST INFRASTRUCTURE ONLY.

``load_reference()`` makes ``import astronomix.<submodule>`` resolve to the files under
/root/reference/astronomix with ``oracle/jaxshim`` standing in for ``jax``, ``jaxtyping``,
``beartype`` and ``equinox`` (none of which is installed in this image).  The package's own
``__init__.py`` is bypassed (it imports ``astropy`` through ``astronomix.units``, which the
finite-volume path never touches): ``astronomix`` is registered as a namespace whose
``__path__`` is the reference directory, so every submodule is the reference's file, executed
as it lies there.  Nothing is copied into this repository.

Only ``oracle/gen_reference_vectors.py`` (run in the build container, where /root/reference
exists) uses this; its outputs are committed under ``tests/golden/reference_vectors/``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AMX_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jaxshim")


def load_reference(x64: bool = True):
    """Return the ``astronomix`` namespace backed by the reference sources on the NumPy shim."""
    pkg_dir = os.path.join(REFERENCE_ROOT, "astronomix")
    if not os.path.isdir(pkg_dir):
        raise FileNotFoundError(f"reference sources not found at {pkg_dir}")
    sys.dont_write_bytecode = True  # /root/reference is read-only and must stay untouched
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import jax  # the stand-in
    if not getattr(jax, "__shim__", False):
        raise RuntimeError("a real jax is importable; use it instead of the NumPy stand-in")
    jax.config.update("jax_enable_x64", x64)
    if "astronomix" not in sys.modules:
        pkg = types.ModuleType("astronomix")
        pkg.__path__ = [pkg_dir]
        pkg.__package__ = "astronomix"
        sys.modules["astronomix"] = pkg
    return sys.modules["astronomix"]


def ref(module: str):
    """Import ``astronomix.<module>`` from the reference tree."""
    load_reference(x64=sys.modules["jax"].config.jax_enable_x64 if "jax" in sys.modules else True)
    return importlib.import_module("astronomix." + module)
