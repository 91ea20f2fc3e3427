"""``beartype`` stand-in: the decorator is the identity (TEST INFRASTRUCTURE ONLY)."""


def beartype(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn
