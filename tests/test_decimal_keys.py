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
