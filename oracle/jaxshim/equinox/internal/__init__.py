from . import _loop  # noqa: F401
