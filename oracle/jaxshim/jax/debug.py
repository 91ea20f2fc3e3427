"""``jax.debug`` stand-in (TEST INFRASTRUCTURE ONLY)."""
import builtins as _builtins


def callback(fn, *args, **kwargs):
    kwargs.pop("ordered", None)
    return fn(*args, **kwargs)


def print(fmt, *args, **kwargs):  # noqa: A001
    kwargs.pop("ordered", None)
    _builtins.print(fmt.format(*args, **kwargs))
