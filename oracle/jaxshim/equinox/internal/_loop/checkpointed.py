"""Forward semantics of equinox's checkpointed_while_loop = lax.while_loop (TEST INFRASTRUCTURE ONLY)."""


def checkpointed_while_loop(cond_fun, body_fun, init_val, *, max_steps=None, checkpoints=None, **_kw):
    val, n = init_val, 0
    while bool(cond_fun(val)) and (max_steps is None or n < max_steps):
        val = body_fun(val)
        n += 1
    return val
