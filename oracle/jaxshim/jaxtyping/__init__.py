"""``jaxtyping`` stand-in: annotations only, no checking (TEST INFRASTRUCTURE ONLY)."""


class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls

    def __getitem__(self, item):
        return self


class Array(_Subscriptable):
    pass


class Float(_Subscriptable):
    pass


class Int(_Subscriptable):
    pass


class Bool(_Subscriptable):
    pass


class Complex(_Subscriptable):
    pass


class PyTree(_Subscriptable):
    pass


class PRNGKeyArray(_Subscriptable):
    pass


def jaxtyped(fn=None, *, typechecker=None):
    if fn is None:
        return lambda f: f
    return fn
