"""The TF-1 API shim (oracle/tf1_shim) against independent numpy / scipy statements of the documented
TensorFlow-1 semantics it stands in for.  The golden vectors are only as good as these ops: NHWC x HWIO
'VALID' cross-correlation without kernel flip, floor division / modulo, gather / gather_nd / scatter
index conventions, value semantics of tensors across variable updates, control flow, TF-1 Adam."""
import os
import sys

import numpy as np
import pytest
from scipy.signal import correlate

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "tf1_shim")


@pytest.fixture()
def tf():
    sys.path.insert(0, SHIM)
    try:
        import tensorflow as tf
        tf.reset_default_graph()
        tf.set_precision("double")
        tf.set_random_seed(1)
        yield tf
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("tensorflow", None)


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_conv_is_valid_cross_correlation_channels_last(tf, nd):
    rng = np.random.default_rng(nd)
    sp, k, ci, co = (7, 6, 5)[:nd], 3, 2, 4
    x = rng.standard_normal((3,) + sp + (ci,))
    w = rng.standard_normal((k,) * nd + (ci, co))
    fn = (lambda a, b: tf.nn.conv1d(a, b, 1, "VALID"), lambda a, b: tf.nn.conv2d(a, b, [1] * 4, "VALID"),
          lambda a, b: tf.nn.conv3d(a, b, [1] * 5, "VALID"))[nd - 1]
    got = fn(tf.constant(x), tf.constant(w)).numpy()
    want = np.zeros((3,) + tuple(s - k + 1 for s in sp) + (co,))
    for n in range(3):
        for o in range(co):
            for i in range(ci):
                want[n, ..., o] += correlate(x[n, ..., i], w[..., i, o], mode="valid")      # no kernel flip
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-12


def test_integer_division_and_modulo_are_floored(tf):
    a = tf.constant(np.array([-7, -1, 0, 5, 7], np.int32))
    assert list((a // 3).numpy()) == [-3, -1, 0, 1, 2]
    assert list((a % 3).numpy()) == [2, 2, 0, 2, 1]
    assert (a - np.int64(2)).numpy().dtype == np.int32          # a Python / numpy operand adopts the tensor's dtype


def test_gather_scatter_conventions_and_value_semantics(tf):
    p = np.arange(24).reshape(4, 6)
    assert np.array_equal(tf.gather(p, tf.constant(np.array([[3, 0], [1, 1]], np.int32))).numpy(), p[[[3, 0], [1, 1]]])
    idx = np.array([[[0, 1], [3, 5]], [[2, 2], [1, 0]]], np.int32)
    assert np.array_equal(tf.gather_nd(tf.constant(p), tf.constant(idx)).numpy(), p[idx[..., 0], idx[..., 1]])
    assert np.array_equal(tf.gather_nd(tf.constant(p), tf.constant(np.array([[[2]], [[0]]], np.int32))).numpy(),
                          p[[[2], [0]]])
    v = tf.Variable(p.astype(np.int32), trainable=False)
    before = v * 1                                   # a tensor computed BEFORE the update keeps the old value
    view = tf.reshape(v, (6, 4))
    tf.scatter_update(v, tf.constant(np.array([2, 0], np.int32)), tf.constant(np.full((2, 6), -1, np.int32)))
    assert np.array_equal(before.numpy(), p) and np.array_equal(view.numpy(), p.reshape(6, 4))
    want = p.copy(); want[[2, 0]] = -1
    assert np.array_equal(v.numpy(), want)
    tf.scatter_nd_update(v, tf.constant(np.array([[1, 1], [3, 5]], np.int32)), tf.constant(np.array([7, 8], np.int32)))
    want[1, 1], want[3, 5] = 7, 8
    assert np.array_equal(v.numpy(), want)
    m = np.array([True, False, True, False])
    assert np.array_equal(tf.boolean_mask(tf.constant(p), tf.constant(m)).numpy(), p[m])


def test_control_flow_and_slicing(tf):
    out = tf.while_loop(lambda i: i < 5, lambda i: i + 2, [tf.constant(0)], parallel_iterations=1, back_prop=False)
    assert int(out.numpy()) == 6
    flag = tf.Variable(True)
    assert tf.cond(flag, lambda: 1, lambda: 2) == 1
    flag.load(False)
    assert tf.cond(flag, lambda: 1, lambda: 2) == 2
    x = np.arange(60).reshape(3, 4, 5)
    assert np.array_equal(tf.slice(tf.constant(x), (0, 1, 2), (-1, 2, -1)).numpy(), x[:, 1:3, 2:])
    assert np.array_equal(tf.tile(tf.constant(x), (1, 3, 2)).numpy(), np.tile(x, (1, 3, 2)))
    assert np.array_equal(tf.transpose(tf.constant(x)).numpy(), x.T)
    assert np.array_equal(tf.map_fn(lambda r: tf.reduce_sum(r, 1), tf.constant(x), back_prop=False).numpy(), x.sum(2))
    assert tf.shape(tf.constant(x))[1:] == (4, 5)


def test_adam_is_the_tf1_update(tf):
    with tf.variable_scope("s"):
        v = tf.get_variable("v", shape=[3], dtype=tf.float32, initializer=tf.constant_initializer(1.0))
    opt = tf.train.AdamOptimizer(1e-2)
    p = np.ones(3); m = np.zeros(3); s = np.zeros(3)
    for t in range(1, 4):
        loss = tf.reduce_sum(v * v * tf.constant(np.array([1.0, 2.0, 3.0])))
        opt.minimize(loss)
        g = 2 * p * np.array([1.0, 2.0, 3.0])
        m = 0.9 * m + 0.1 * g
        s = 0.999 * s + 0.001 * g * g
        p = p - 1e-2 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(s) + 1e-8)     # epsilon outside the root
        assert np.abs(v.numpy() - p).max() < 1e-12


def test_precision_modes_and_logged_randomness(tf):
    tf.set_precision("single")
    x = tf.cast(tf.constant(np.array([1, 2], np.int32)), tf.float32)
    assert x.numpy().dtype == np.float32 and tf.complex(x, x).numpy().dtype == np.complex64
    tf.set_precision("double")
    assert tf.cast(x, tf.float32).numpy().dtype == np.float64
    tf.set_random_seed(5)
    a = tf.random_uniform([4, 3], 0, 9, dtype=tf.int32).numpy()
    u = tf.random_uniform([5], 0., 1., dtype=tf.float32).numpy()
    assert a.min() >= 0 and a.max() < 9 and 0 <= u.min() and u.max() < 1
    assert [k for k, _ in tf.random_log] == ["uniform_int", "uniform_float"]
    tf.set_random_seed(5)
    assert np.array_equal(tf.random_uniform([4, 3], 0, 9, dtype=tf.int32).numpy(), a)
