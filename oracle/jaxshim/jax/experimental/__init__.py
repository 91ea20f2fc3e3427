"""``jax.experimental`` stand-in (TEST INFRASTRUCTURE ONLY)."""
from . import checkify  # noqa: F401
