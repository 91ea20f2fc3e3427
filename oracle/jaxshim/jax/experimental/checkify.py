"""``jax.experimental.checkify`` stand-in: checks raise immediately (TEST INFRASTRUCTURE ONLY)."""

user_checks = index_checks = float_checks = nan_checks = div_checks = frozenset()


class _NoError:
    def throw(self):
        return None


def checkify(fun, errors=None):
    def checked(*args, **kwargs):
        return _NoError(), fun(*args, **kwargs)
    return checked


def check(pred, msg, *args, **kwargs):
    if not bool(pred):
        raise ValueError(msg.format(*args, **kwargs))
