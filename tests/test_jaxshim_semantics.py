"""The JAX rules that oracle/jaxshim has to reproduce for the recorded reference vectors to mean
what they claim (oracle/jaxshim/README.md).  Run in a child process: the stand-in registers modules
named ``jax`` / ``equinox`` / ``jaxtyping`` / ``beartype``, which must not leak into this one."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECKS = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import jax, jax.numpy as jnp
assert getattr(jax, "__shim__", False)

# 1. functional updates: a copy is updated, the operand is untouched
x = jnp.arange(6.0).reshape(2, 3)
y = x.at[0, 1].set(9.0)
assert float(x[0, 1]) == 1.0 and float(y[0, 1]) == 9.0
z = x.at[1].add(1.0)
assert np.array_equal(np.asarray(z[1]), [4.0, 5.0, 6.0]) and np.array_equal(np.asarray(x[1]), [3.0, 4.0, 5.0])
w = x.at[:, 0:2].set(jnp.array([7.0, 8.0])[None, :])
assert np.array_equal(np.asarray(w[:, :2]), [[7.0, 8.0], [7.0, 8.0]])

# 2. out-of-range static integer index: reads clamp, scatter updates are dropped (jax's default modes)
v = jnp.arange(3.0)
assert float(v[5]) == 2.0
assert np.array_equal(np.asarray(v.at[3].set(-1.0)), [0.0, 1.0, 2.0])
m = jnp.arange(12.0).reshape(3, 4)
assert np.array_equal(np.asarray(m.at[3, 1].set(-1.0)), np.asarray(m))
assert np.array_equal(np.asarray(m[3, :]), np.asarray(m[2, :]))
assert np.array_equal(np.asarray(v.at[-1].set(5.0)), [0.0, 1.0, 5.0])  # negative in-range indices wrap as in NumPy

# 3. dtypes: x64 flag decides the defaults; Python scalars are weak; no 64-bit type survives with x64 off
jax.config.update("jax_enable_x64", False)
a = jnp.zeros((4,))
assert a.dtype == np.float32
assert (0.5 * a + 1.0 / 3.0).dtype == np.float32 and (a * 2).dtype == np.float32
assert jnp.array([5 / 3, 800.0]).dtype == np.float32 and jnp.linspace(0, 1, 5).dtype == np.float32
assert jnp.sqrt(a + 2.0).dtype == np.float32 and jnp.max(a).dtype == np.float32
assert jnp.where(a > 0, a, 1.0).dtype == np.float32
assert (a.at[0].set(1 / 3))[0] == np.float32(1 / 3)
jax.config.update("jax_enable_x64", True)
assert jnp.zeros((4,)).dtype == np.float64 and jnp.array([5 / 3]).dtype == np.float64

# control flow: while_loop / fori_loop / cond carry Python-level semantics of the jax primitives
n, acc = jax.lax.while_loop(lambda s: s[0] < 5, lambda s: (s[0] + 1, s[1] + s[0]), (0, 0))
assert (n, acc) == (5, 10)
assert jax.lax.fori_loop(0, 4, lambda i, s: s + i, 0) == 6
assert jax.lax.cond(True, lambda u: u + 1, lambda u: u - 1, 1) == 2
assert float(jax.lax.cond(jnp.array(False), lambda: 1.0, lambda: 0.0)) == 0.0

# jit with static arguments is the identity call
from functools import partial
@partial(jax.jit, static_argnames=["k"])
def f(u, k):
    return u * k
assert float(f(jnp.array(2.0), k=3)) == 6.0

# the stencil primitives the path is built from
r = jnp.arange(5.0)
assert np.array_equal(np.asarray(jnp.roll(r, 1)), [4.0, 0.0, 1.0, 2.0, 3.0])
assert abs(float(jnp.linalg.norm(jnp.array([[3.0], [4.0]]), axis=0)[0]) - 5.0) < 1e-15
print("ok")
'''


def test_stand_in_follows_the_jax_rules_the_reference_relies_on():
    shim = os.path.join(ROOT, "oracle", "jaxshim")
    res = subprocess.run([sys.executable, "-c", CHECKS % shim], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout + res.stderr


def test_stand_in_does_not_leak_into_this_process():
    for name in ("jax", "equinox"):
        mod = sys.modules.get(name)
        assert mod is None or not getattr(mod, "__shim__", False)
