"""jax's default PRNG, restated (oracle/threefry.py) and on the device (amx_threefry_normal).

Pins: the three Random123 known-answer vectors of Threefry-2x32-20 (Salmon et al. SC'11; the same three are in
jax's own tests/random_test.py::testThreefry2x32), and values the jax documentation prints -- old counter layout
(jax < 0.5 default): random.normal(PRNGKey(42)) = -0.18471177, random.normal(PRNGKey(0), (1,)) = -0.20584226,
random.uniform(PRNGKey(0)) = 0.41845703; partitionable layout (jax >= 0.5 default): random.uniform(key(0)) =
0.947667, random.normal(key(0)) = 1.6226422.  (Quoted from the published documentation; jax cannot be run here.)"""
import numpy as np
import pytest

from astronomix_b200 import threefry as tfh
from oracle import threefry as tf


def test_random123_known_answers():
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for k, c, want in kat:
        o0, o1 = tf.threefry2x32(k[0], k[1], [c[0]], [c[1]])
        assert (int(o0[0]), int(o1[0])) == want
        assert tfh._block(k[0], k[1], c[0], c[1]) == want  # the host-side key arithmetic of the product


def test_documented_jax_values_both_layouts():
    k42, k0 = tf.seed_key(42), tf.seed_key(0)
    assert abs(float(tf.normal(k42, (), np.float32, partitionable=False)) - (-0.18471177)) < 5e-8
    assert abs(float(tf.normal(k0, (1,), np.float32, partitionable=False)[0]) - (-0.20584226)) < 5e-8
    assert abs(float(tf.uniform(k0, (), np.float32, 0.0, 1.0, partitionable=False)) - 0.41845703) < 5e-8
    assert abs(float(tf.uniform(k0, (), np.float32, 0.0, 1.0, partitionable=True)) - 0.947667) < 5e-7
    assert abs(float(tf.normal(k0, (), np.float32, partitionable=True)) - 1.6226422) < 2e-7


@pytest.mark.parametrize("part", [True, False])
def test_host_split_equals_restatement(part):
    k = tf.seed_key(42)
    for _ in range(3):
        a = tf.split(k, 3, part)
        b = tfh.split((int(k[0]), int(k[1])), 3, part)
        assert [(int(x), int(y)) for x, y in a] == b
        k = a[0]


def test_normal_statistics():
    x = tf.normal(tf.seed_key(7), (3, 32, 32, 32), np.float32)
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1.0) < 0.02
    y = tf.normal(tf.seed_key(7), (3, 32, 32, 32), np.float64)
    assert abs(y.mean()) < 0.02 and abs(y.std() - 1.0) < 0.02
    # fp64 draws use the high and low words of the same blocks: not the fp32 stream widened
    assert not np.allclose(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("part", [True, False])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(3, 16, 16, 16), (5,), (3, 7, 11)])
def test_device_normal_equals_restatement(part, dtype, shape):
    import torch
    k = tf.seed_key(42)
    sk = tf.split(k, 3, part)[1]
    want = tf.normal(sk, shape, dtype, part)
    got = tfh.normal((int(sk[0]), int(sk[1])), shape, torch.float32 if dtype == np.float32 else torch.float64,
                     partitionable=part).cpu().numpy()
    # same bits -> same uniforms; erf_inv differs by the log1p / polynomial rounding of the two libraries
    tol = 3e-6 if dtype == np.float32 else 1e-12
    assert np.max(np.abs(got - want) / (1.0 + np.abs(want))) < tol
