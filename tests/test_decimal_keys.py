"""The flood fill (History::bfs, history.cpp:96-137) identifies a cell by std::to_string of its coordinates,
i.e. by "%f": six decimals, correctly rounded, ties to even.  The kernel computes that integer exactly
(csrc/frontier_kernels.cuh: dec6_magnitude); this checks the same host+device function on the CPU against the
C library's own formatting (Python's '%f' is the same correctly rounded conversion)."""
import ctypes as C
import math

import numpy as np
import pytest

from nbv_exploration_planner_b200 import _native as N


def _dec6(x: float):
    out = C.c_uint64(0)
    rc = N.lib().nbvgpu_debug_round_decimal6(C.c_double(x), C.byref(out))
    return rc, out.value


def _want(x: float) -> int:
    s = "%f" % abs(x)
    return int(s.replace(".", ""))


def test_matches_printf_on_random_and_lattice_values():
    rng = np.random.default_rng(3)
    vs = float(np.float32(0.15))
    xs = np.concatenate([rng.uniform(-1000, 1000, 20000), rng.uniform(-1e-5, 1e-5, 5000), rng.uniform(-3e5, 3e5, 5000),
                         10.03 + vs * rng.integers(-60, 60, 5000), np.float64(np.float32(0.2)) * rng.integers(-5000, 5000, 5000),
                         [0.0, -0.0, 1e-7, -1e-7, 4.9e-7, 5.1e-7, 1e-300, 5e-324, 1048575.999999]])
    for x in xs:
        rc, got = _dec6(float(x))
        assert rc == 0 and got == _want(float(x)), (x, got, _want(float(x)))


def test_exact_ties_round_to_even():
    """x * 10^6 is exactly a half-integer only for dyadic x = (2j+1) * 15625 / 2^(7+k) ...: 2^-7 = 0.0078125 ->
    7812.5 -> 7812 (even); 3 * 2^-7 = 0.0234375 -> 23437.5 -> 23438 (even)."""
    for x, want in ((2.0 ** -7, 7812), (3 * 2.0 ** -7, 23438), (5 * 2.0 ** -7, 39062), (1 + 2.0 ** -7, 1007812), (-(2.0 ** -7), 7812),
                    (2.0 ** -6 + 2.0 ** -7, 23438), (100 + 7 * 2.0 ** -7, 100054688)):
        assert abs(x) * 1e6 % 1 == 0.5          # really a tie (exactly representable product)
        rc, got = _dec6(x)
        assert rc == 0 and got == want == _want(x), (x, got, want, _want(x))
    # one ulp either side of a tie goes to the nearer integer
    t = 1 + 2.0 ** -7
    assert _dec6(math.nextafter(t, 2.0))[1] == 1007813 == _want(math.nextafter(t, 2.0))
    assert _dec6(math.nextafter(t, 0.0))[1] == 1007812 == _want(math.nextafter(t, 0.0))


def test_out_of_range_is_reported():
    for x in (float("inf"), float("nan"), 2.0 ** 21, -3e6):
        rc, _ = _dec6(x)
        assert rc != 0


def test_squared_distance_threshold_is_equivalent_to_the_norm_test():
    """The flood fill compares the squared distance with the largest double whose correctly rounded square root does not
    exceed max_distance instead of taking the root (history.cpp:118: (candidate - point).norm() <= max_distance).  Around
    the threshold, and on random sums, both tests must agree for every value."""
    rng = np.random.default_rng(9)
    ds = np.concatenate([[0.0, 6.0, 1.0, 0.5, 2.0 ** 0.5, 1e-3, 123.456, 4499.0],
                         rng.uniform(0.01, 50.0, 300), np.float64(np.float32(rng.uniform(0.1, 10.0, 100)))])
    for d in ds:
        t = C.c_double(0.0)
        assert N.lib().nbvgpu_debug_sqrt_threshold(C.c_double(d), C.byref(t)) == N.OK
        t = t.value
        assert math.sqrt(t) <= d < math.sqrt(np.nextafter(t, np.inf))
        x = t
        for _ in range(40):                          # 40 doubles below ...
            assert math.sqrt(x) <= d
            x = float(np.nextafter(x, -np.inf))
            if x < 0:
                break
        x = float(np.nextafter(t, np.inf))
        for _ in range(40):                          # ... and above the threshold
            assert not (math.sqrt(x) <= d)
            x = float(np.nextafter(x, np.inf))
        xs = rng.uniform(0.0, 2.0 * d * d + 1e-9, 2000)
        assert np.array_equal(np.sqrt(xs) <= d, xs <= t)
    bad = C.c_double(0.0)
    for d in (-1.0, float("nan"), float("inf")):
        assert N.lib().nbvgpu_debug_sqrt_threshold(C.c_double(d), C.byref(bad)) == N.ERR_INVALID
