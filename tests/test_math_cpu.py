"""CPU-only accuracy tests of templates/hb_math.cuh (the same source the kernels compile, built
here as host C++ with HB_HOST_TEST) against libm."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "hydromodels.jl_b200", "csrc", "templates", "hb_math.cuh")


def _build(extra_define=""):
    d = tempfile.mkdtemp(prefix="hb_math_")
    src = os.path.join(d, "m.cpp")
    open(src, "w").write(f'''#define HB_HOST_TEST 1
{extra_define}
#include "{HDR}"
extern "C" void apply(int which, long n, const double* x, double* y) {{
    for (long i = 0; i < n; ++i) {{
        switch (which) {{
            case 0: y[i] = hb_exp(x[i]); break;
            case 1: y[i] = hb_step(x[i]); break;
            case 2: y[i] = hb_tanh(x[i]); break;
            case 3: y[i] = hb_rcp(x[i]); break;
            case 4: y[i] = hb_log_pos(x[i]); break;
        }}
    }}
}}
extern "C" void apply2(long n, const double* x, const double* y, double* z) {{
    for (long i = 0; i < n; ++i) z[i] = hb_pow(x[i], y[i]);
}}''')
    so = os.path.join(d, "m.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.apply2.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]

    def f(which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib.apply(C.c_int(which), C.c_long(x.size), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
        return y

    def pw(x, yy):
        x = np.ascontiguousarray(x, dtype=np.float64)
        yy = np.ascontiguousarray(yy, dtype=np.float64)
        z = np.empty_like(x)
        lib.apply2(C.c_long(x.size), x.ctypes.data_as(C.c_void_p), yy.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p))
        return z
    f.pow = pw
    return f


@pytest.fixture(scope="module")
def hbm():
    return _build()


@pytest.fixture(scope="module")
def hbm_tanh10():
    return _build("#define HB_TANH_EXACT 0")


@pytest.fixture(scope="module")
def hbm_nan():
    return _build("#define HB_NAN_PROPAGATE 1")


def test_exp_full_range(hbm):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-2, 2, 200000), rng.normal(0, 1e-3, 1000),
                        np.array([0.0, -0.0, 1.0, -1.0, 709.7, -708.0, 1e-300])])
    rel = np.abs(hbm(0, x) - np.exp(x)) / np.exp(x)
    assert rel.max() < 5e-16          # < 2.3 ulp (table entry, polynomial and product each round once)
    edge = hbm(0, np.array([710.0, 1e6, np.inf, -746.0, -1e6, -np.inf, -740.0, -745.0]))
    assert np.all(np.isinf(edge[:3])) and np.all(edge[3:6] == 0.0)
    assert np.allclose(edge[6:], np.exp([-740.0, -745.0]), rtol=0, atol=1e-323)  # subnormal results, gradual underflow


def test_step_func_against_reference_form(hbm):
    """|hb_step(x) − (tanh(5x)+1)/2| ≤ 2e-16 absolute everywhere (the reference form itself only
    resolves to ~1.1e-16 absolute), and relative 1e-15 where the value is not tiny."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-100, 100, 200000), rng.uniform(-3, 3, 200000), rng.normal(0, 1e-6, 1000),
                        np.array([0.0, 1e-6, 1303.0, -1303.0, 1e300, -1e300, np.inf, -np.inf])])
    got = hbm(1, x)
    ref = (np.tanh(5.0 * x) + 1.0) * 0.5
    assert np.all(np.isfinite(got)) and got.min() >= 0.0 and got.max() <= 1.0
    assert np.abs(got - ref).max() < 2.5e-16
    # against the exact logistic in extended precision the relative error is ~1 ulp even where the
    # value is tiny (the reference's tanh form loses all relative accuracy there)
    xl = np.clip(x, -60.0, 60.0).astype(np.longdouble)
    truth = 1.0 / (1.0 + np.exp(-10.0 * xl))
    sel = np.abs(x) <= 60.0
    # conditioning: rounding z = -10x to double alone moves exp(z) by |z|·1.1e-16 relative
    tol = 6e-16 + np.abs(10.0 * x[sel]) * 1.2e-16
    assert np.all(np.abs(got[sel] - truth[sel]) / truth[sel] < tol)


def test_tanh_and_rcp(hbm, hbm_tanh10):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-30, 30, 200000), rng.uniform(-3, 3, 400000), rng.normal(0, 1e-4, 1000), [0.0, 400.0, -400.0]])
    assert np.abs(hbm(2, x) - np.tanh(x)).max() < 4.5e-16      # 2 ulp of 1.0
    # opt-in form with 10 FP64-pipe operations (the tail of e^(2r) in float): 1e-13 absolute
    assert np.abs(hbm_tanh10(2, x) - np.tanh(x)).max() < 1e-13
    y = np.concatenate([rng.uniform(1e-3, 1e3, 100000), -rng.uniform(1e-3, 1e3, 100000), 10.0 ** rng.uniform(-250, 250, 100000)])
    assert (np.abs(hbm(3, y) * y - 1.0)).max() < 5e-16


@pytest.fixture(scope="module")
def hbm_logpoly():
    return _build("#define HB_LOG_TABLE 0")


def test_table_log_and_pow(hbm):
    """Table form of hb_log_pos (the default wherever a body raises to a power): |error| <= 7e-16 max(|log x|, 1), and
    hb_pow = exp(y log x) within the rounding of the product plus |y| times that."""
    rng = np.random.default_rng(3)
    x = np.abs(np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 200000), rng.uniform(1e-9, 1.0, 200000),
                               1.0 + rng.normal(0, 1e-6, 2000), [1.0, 2.0, 0.5, np.sqrt(2.0), 1.0 - 2.0 ** -53, 1.0 + 2.0 ** -52]]))
    ref = np.log(x.astype(np.longdouble))
    err = np.abs(hbm(4, x) - ref).astype(np.float64)
    assert np.all(err <= 7e-16 * np.maximum(np.abs(ref).astype(np.float64), 1.0))
    xb = np.concatenate([rng.uniform(0.0, 1.0, 200000), rng.uniform(1e-12, 1e3, 100000)])
    yb = np.concatenate([rng.uniform(0.05, 8.0, 200000), rng.uniform(-3.0, 6.0, 100000)])
    got, refp = hbm.pow(xb, yb), np.power(xb.astype(np.longdouble), yb.astype(np.longdouble))
    ok = (refp > 1e-300) & np.isfinite(refp)
    ylx = np.abs(yb * np.log(np.maximum(xb, 1e-300)))
    bound = 3.0 * 1.12e-16 * (1.0 + ylx) + 8e-16 * np.abs(yb) * np.maximum(np.abs(np.log(np.maximum(xb, 1e-300))), 1.0)
    assert np.all(np.abs((got[ok] - refp[ok]) / refp[ok]).astype(np.float64) <= bound[ok])
    assert hbm.pow([0.0, 0.0, 0.0, 5.0, 1e-320], [2.5, 0.0, 1.0, 0.0, 3.0]).tolist() == [0.0, 1.0, 0.0, 1.0, 0.0]
    assert np.isnan(hbm.pow([-1.0], [2.5])[0]) and np.isinf(hbm.pow([0.0], [-1.0])[0])


def test_log_and_pow(hbm_logpoly):
    hbm = hbm_logpoly
    """hb_log_pos: < 1 ulp-ish against libm over the normal range; hb_pow = exp(y log x): relative error bounded by the rounding
    of the product y*log(x) (~|y log x| * 1.2e-16) plus the two kernels, and the x = 0 / y = 0 limits of the reference's `^`."""
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 200000), 1.0 + rng.normal(0, 1e-6, 2000), [1.0, 2.0, 0.5, np.sqrt(2.0)]])
    x = np.abs(x)
    err = np.abs(hbm(4, x) - np.log(x))
    assert np.all(err <= 2.3e-16 * np.maximum(np.abs(np.log(x)), 1e-300) + 1.2e-16 * np.abs(x - 1.0) + 1e-300)
    xb = np.concatenate([rng.uniform(0.0, 1.0, 200000), rng.uniform(1e-12, 1e3, 100000)])
    yb = np.concatenate([rng.uniform(0.05, 8.0, 200000), rng.uniform(-3.0, 6.0, 100000)])
    got, ref = hbm.pow(xb, yb), np.power(xb.astype(np.longdouble), yb.astype(np.longdouble))
    ok = (ref > 1e-300) & np.isfinite(ref)
    bound = 3.0 * 1.12e-16 * (1.0 + np.abs(yb * np.log(np.maximum(xb, 1e-300))))      # measured: 2.5 eps (1 + |y log x|), worst 8.2e-15
    assert np.all(np.abs((got[ok] - ref[ok]) / ref[ok]).astype(np.float64) <= bound[ok])
    assert hbm.pow([0.0, 0.0, 0.0, 5.0, 1e-320], [2.5, 0.0, 1.0, 0.0, 3.0]).tolist() == [0.0, 1.0, 0.0, 1.0, 0.0]
    assert np.isnan(hbm.pow([-1.0], [2.5])[0]) and np.isinf(hbm.pow([0.0], [-1.0])[0])


def test_nan_propagates_and_infinities_saturate(hbm_nan):
    """With `HB_NAN_PROPAGATE 1` NaN forcing does not come out of exp / step_func / tanh as a plausible finite number (Julia
    propagates it); +/-inf keep their limits either way."""
    hbm = hbm_nan
    nan, inf = float("nan"), float("inf")
    for which in (0, 1, 2):
        assert np.isnan(hbm(which, [nan, -nan])).all()
    assert hbm(0, [inf])[0] == inf and hbm(0, [-inf])[0] == 0.0
    assert hbm(1, [inf])[0] == 1.0 and hbm(1, [-inf])[0] < 1e-300
    assert hbm(2, [inf])[0] == 1.0 and hbm(2, [-inf])[0] == -1.0
