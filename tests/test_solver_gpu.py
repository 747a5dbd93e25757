"""The small dense solver of the reduced camera system (what Ceres hands to LAPACK / CHOLMOD, CeresOptimizer.cc:178-187,
516-519) on its own: cmos_debug_solve_spd runs exactly the device routines k_solve_small / k_cr_factor use
(factor_and_invert24 + back_substitute24) on a caller-supplied SPD matrix; numpy's LAPACK solve is the checker.

Sizes cover both code paths (register-resident tiles up to 152 unknowns, shared-memory panels above), every residue of n
modulo 8 and modulo 24 that a multiple of 6 can have, and the sizes the BA tests actually factor (114, 120, 144)."""
import ctypes as C

import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import _lib

pytestmark = pytest.mark.gpu

TOL = 1e-11          # relative to |x|, for matrices of condition <= 1e3 (fp64: ~1e-13 expected)


def _spd(n, seed, banded=False):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    if banded:                                   # block-banded like a covisibility window, still dense inside the band
        i, j = np.indices((n, n))
        B[np.abs(i - j) > 30] = 0.0
    A = B @ B.T / n + np.eye(n)
    d = 1.0 / np.sqrt(np.diag(A))                # Jacobi-scaled like the LM system
    return (A * d[:, None]) * d[None, :]


def _solve(A, b, variant=0):
    n = A.shape[0]
    x = np.zeros(n); failed = C.c_int32(-1); cyc = (C.c_int64 * 2)()
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
    _lib.check(_lib.lib().cmos_debug_solve_spd(_lib.ptr(A), _lib.ptr(b), n, variant, _lib.ptr(x), C.byref(failed), cyc))
    return x, failed.value, (cyc[0], cyc[1])


@pytest.mark.parametrize("n", [6, 12, 18, 24, 30, 48, 54, 60, 96, 102, 114, 120, 126, 132, 144, 150, 156, 180, 228])
def test_solve_equals_lapack(n):
    for seed, banded in ((n, False), (1000 + n, True)):
        A = _spd(n, seed, banded)
        b = np.random.default_rng(7 * n + seed).standard_normal(n)
        ref = np.linalg.solve(A, b)
        for variant in (0, 1):
            x, failed, cyc = _solve(A, b, variant)
            assert failed == 0
            err = np.abs(x - ref).max() / np.abs(ref).max()
            assert err < TOL, (n, banded, variant, err)
            assert cyc[0] > 0


def test_identity_and_diagonal():
    for n in (24, 114, 120):
        d = np.linspace(0.5, 4.0, n)
        b = np.arange(1.0, n + 1.0)
        x, failed, _ = _solve(np.diag(d), b)
        assert failed == 0 and np.allclose(x, b / d, rtol=1e-14, atol=0)


@pytest.mark.parametrize("n", [24, 114, 120, 150, 180])
def test_indefinite_matrix_is_reported(n):
    A = _spd(n, 5)
    A[n // 2, n // 2] = -1.0                      # a negative pivot appears at that column
    for variant in (0, 1):
        assert _solve(A, np.ones(n), variant)[1] == 1
    A = _spd(n, 6)
    A[n - 1, n - 1] = np.nan
    _, failed, _ = _solve(A, np.ones(n))
    assert failed == 1


def test_repeatable_bitwise():
    A = _spd(120, 9); b = np.ones(120)
    x0, _, _ = _solve(A, b)
    x1, _, _ = _solve(A, b)
    assert np.array_equal(x0, x1)
