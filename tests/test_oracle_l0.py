"""CPU tests of the oracle's L0 layer (oracle/l0.c): the stand-in for the reference's
absent RNG library.  Special functions are pinned against scipy.special on the
argument ranges the samplers generate; the Philox block
function against the Random123 known-answer vectors; the variate generators
against their distributions."""
import ctypes as C

import numpy as np
import pytest
from scipy import special, stats


@pytest.fixture(scope="module")
def l0(port):
    lib = port.lib
    lib.pgo_p_norm.argtypes = [C.c_double, C.c_int]
    lib.pgo_p_norm.restype = C.c_double
    lib.pgo_p_gamma_rate.argtypes = [C.c_double] * 3
    lib.pgo_p_gamma_rate.restype = C.c_double
    lib.pgo_p_igauss.argtypes = [C.c_double] * 3
    lib.pgo_p_igauss.restype = C.c_double
    lib.pgo_Gamma.argtypes = [C.c_double, C.c_int]
    lib.pgo_Gamma.restype = C.c_double
    lib.pgo_philox4x32_10.argtypes = [C.c_void_p] * 3
    return lib


def test_philox_known_answers(l0):
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        l0.pgo_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert out.tolist() == want


def test_p_norm(l0):
    xs = np.concatenate([np.linspace(-40, 8, 977), [-1e3, -300.0, -38.0, -37.0, -20.0, -19.99, 0.0]])
    for x in xs:
        got = l0.pgo_p_norm(x, 0)
        want = special.ndtr(x)
        # the condition number of Phi in the lower tail is x^2
        assert got == pytest.approx(want, rel=1e-15 * max(100.0, 2 * x * x), abs=1e-300)
        gl = l0.pgo_p_norm(x, 1)
        wl = special.log_ndtr(x)
        assert gl == pytest.approx(wl, rel=2e-13, abs=1e-16)


def test_p_gamma_rate(l0):
    rng = np.random.default_rng(0)
    # Alt: shape 1/2 with x*rate in [0.78,1.94]; shape h in [1,4]; SP: shape n in (13,170]
    cases = [(0.5, rng.uniform(0.5, 3, 300)), (1.0, rng.uniform(0.5, 60, 300))]
    for a in rng.uniform(1, 4, 20):
        cases.append((a, rng.uniform(0.3, 80, 50)))
    for a in rng.uniform(13, 170, 40):
        cases.append((a, a * rng.uniform(0.2, 3.0, 50)))
    for a, xs in cases:
        for x in xs:
            got = l0.pgo_p_gamma_rate(x, a, 1.0)
            want = special.gammainc(a, x)  # scipy itself is only good to ~1e-14 here
            assert abs(got - want) <= 2e-14, (a, x, got, want)
    # tighter pin against arbitrary precision where mpmath is installed
    try:
        import mpmath
    except ImportError:
        mpmath = None
    if mpmath is not None:
        mpmath.mp.dps = 40
        for a, xs in cases[::7]:
            for x in xs[::10]:
                got = l0.pgo_p_gamma_rate(x, a, 1.0)
                want = float(mpmath.gammainc(a, 0, x, regularized=True))
                assert abs(got - want) <= 1.5e-15, (a, x, got, want)
    # rate scaling: P(shape, x*rate)
    assert l0.pgo_p_gamma_rate(2.0, 3.0, 1.5) == pytest.approx(special.gammainc(3.0, 3.0), rel=1e-14)


def test_p_igauss(l0):
    rng = np.random.default_rng(1)
    for _ in range(400):
        x = rng.uniform(0.05, 1.1)
        mu = rng.uniform(0.05, 5)
        lam = rng.uniform(1, 170)
        got = l0.pgo_p_igauss(x, mu, lam)
        want = stats.invgauss.cdf(x, mu / lam, scale=lam)
        assert got == pytest.approx(want, rel=1e-9, abs=1e-300)
    # log-space form must survive 2*lambda/mu >> 709
    v = l0.pgo_p_igauss(0.9, 0.2, 150.0)
    assert np.isfinite(v) and 0.0 < v <= 1.0


def test_gamma_function(l0):
    for x in np.linspace(1, 170, 200):
        assert l0.pgo_Gamma(x, 1) == pytest.approx(special.gammaln(x), rel=1e-14, abs=1e-15)
        assert l0.pgo_Gamma(x, 0) == pytest.approx(special.gamma(x), rel=1e-13)


def _draw_many(port, fn, n, seed):
    """Run a composite generator n times on independent Philox streams."""
    import ctypes as C
    from oracle.loader import Stream  # noqa: F401
    lib = port.lib

    class Src(C.Structure):
        _fields_ = [("raw", C.c_byte * 256)]
    lib.pgo_src_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
    out = np.empty(n)
    s = Src()
    for i in range(n):
        lib.pgo_src_philox(C.byref(s), seed, i, 0)
        out[i] = fn(C.byref(s))
    return out


def test_primitive_variates(port):
    lib = port.lib
    for f in (lib.pgo_unif, lib.pgo_expon, lib.pgo_norm):
        f.argtypes = [C.c_void_p]
        f.restype = C.c_double
    lib.pgo_gamma.argtypes = [C.c_void_p, C.c_double]
    lib.pgo_gamma.restype = C.c_double
    n = 20000
    assert stats.kstest(_draw_many(port, lib.pgo_unif, n, 1), "uniform").pvalue > 1e-3
    assert stats.kstest(_draw_many(port, lib.pgo_expon, n, 2), "expon").pvalue > 1e-3
    assert stats.kstest(_draw_many(port, lib.pgo_norm, n, 3), "norm").pvalue > 1e-3
    for a in (0.3, 1.0, 2.5, 40.0):
        x = _draw_many(port, lambda s: lib.pgo_gamma(s, a), n, 4)
        assert stats.kstest(x, "gamma", args=(a,)).pvalue > 1e-3


def test_composite_variates(port):
    lib = port.lib
    lib.pgo_igauss.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.pgo_igauss.restype = C.c_double
    lib.pgo_ltgamma.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    lib.pgo_ltgamma.restype = C.c_double
    lib.pgo_rtinvchi2.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.pgo_rtinvchi2.restype = C.c_double
    lib.pgo_tnorm.argtypes = [C.c_void_p] + [C.c_double] * 4
    lib.pgo_tnorm.restype = C.c_double
    n = 20000
    mu, lam = 0.7, 3.0
    x = _draw_many(port, lambda s: lib.pgo_igauss(s, mu, lam), n, 5)
    assert stats.kstest(x, "invgauss", args=(mu / lam, 0, lam)).pvalue > 1e-3
    # left-truncated gamma: compare with the conditional CDF
    for shape, rate, tr in ((1.0, 2.0, 0.64), (2.5, 1.7, 1.2), (40.0, 60.0, 0.9)):
        x = _draw_many(port, lambda s: lib.pgo_ltgamma(s, shape, rate, tr), n, 6)
        assert x.min() >= tr
        g = stats.gamma(shape, scale=1 / rate)
        cdf = lambda v: (g.cdf(v) - g.cdf(tr)) / g.sf(tr)
        assert stats.kstest(x, cdf).pvalue > 1e-3
    # right-truncated scaled inverse chi^2(1): X = scale/Z^2, Z|Z > 1/sqrt(trunc/scale)
    scale, tr = 20.0, 0.9
    x = _draw_many(port, lambda s: lib.pgo_rtinvchi2(s, scale, tr), n, 7)
    assert x.max() <= tr
    left = 1 / np.sqrt(tr / scale)
    cdf = lambda v: stats.norm.sf(np.sqrt(scale / v)) / stats.norm.sf(left)
    assert stats.kstest(x, cdf).pvalue > 1e-3
    # two-sided truncated normal
    # (far tails use Robert's rejection samplers, the rest an inverse CDF)
    for lo, hi in ((-0.5, 1.5), (2.0, np.inf), (-np.inf, -3.0), (4.0, 4.5), (6.0, 6.05), (5.0, np.inf),
                   (-np.inf, -7.0), (-300.0, -250.0), (40.0, 41.0), (-3.0, 60.0)):
        x = _draw_many(port, lambda s: lib.pgo_tnorm(s, lo, hi, 0.0, 1.0), 5000, 8)
        assert np.all(np.isfinite(x)) and x.min() >= lo and x.max() <= hi
        if lo >= 0:      # evaluate the conditional CDF on the tail that keeps precision
            cdf = lambda v: (stats.norm.logsf(lo) - 0 > -np.inf) * \
                (1 - np.exp(stats.norm.logsf(v) - stats.norm.logsf(lo))) / \
                (1 - np.exp(stats.norm.logsf(hi) - stats.norm.logsf(lo)))
        elif hi <= 0:
            cdf = lambda v: (np.exp(stats.norm.logcdf(v) - stats.norm.logcdf(hi)) -
                             np.exp(stats.norm.logcdf(lo) - stats.norm.logcdf(hi))) / \
                (1 - np.exp(stats.norm.logcdf(lo) - stats.norm.logcdf(hi)))
        else:
            cdf = lambda v: (stats.norm.cdf(v) - stats.norm.cdf(lo)) / (stats.norm.cdf(hi) - stats.norm.cdf(lo))
        assert stats.kstest(x, cdf).pvalue > 1e-3, (lo, hi)
    # with a location and scale, and a degenerate interval
    x = _draw_many(port, lambda s: lib.pgo_tnorm(s, 1.0, 2.0, 5.0, 0.5), 4000, 9)
    assert x.min() >= 1.0 and x.max() <= 2.0 and x.mean() > 1.8
    assert lib.pgo_tnorm(C.byref((C.c_byte * 256)()), 2.0, 2.0, 0.0, 1.0) == 2.0
