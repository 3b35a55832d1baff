"""Assertion helpers mirroring the reference's test vocabulary (lib/src/lib.rs:4-5, approx 0.3.2)."""
import numpy as np

F32_EPSILON = np.float32(1.1920929e-7)


def f(x):
    return np.asarray(x, dtype=np.float32)


def assert_eq(actual, expected, msg=""):
    """Rust assert_eq! on f32 values: exact after rounding the literal to f32."""
    a, e = f(actual), f(expected)
    assert a.shape == e.shape and np.array_equal(a, e), f"{msg} expected {e!r}, got {a!r}"


def assert_abs_diff_eq(actual, expected, epsilon=F32_EPSILON, msg=""):
    """approx::assert_abs_diff_eq! with the f32 default epsilon (tuple.rs:161-163, color.rs:82-84)."""
    a, e = f(actual), f(expected)
    d = np.abs(a - e)
    assert a.shape == e.shape and np.all(d <= np.float32(epsilon)), f"{msg} expected {e!r}, got {a!r} (|d|={d!r})"
