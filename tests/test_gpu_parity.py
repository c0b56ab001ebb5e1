"""GPU parity tests: the CUDA engine, called through its C ABI, against the oracle.

Tier 1 (north_star): fed the identical injected uniform/exponential/normal(/gamma)
variate stream, every accept/reject decision matches (identical variate
consumption per observation) and draws agree within REL = 1e-12 relative.
Stream contract: from the same Philox key/counter the engine and the oracle's
independent C restatement produce the same draws (same bar), at sizes the oracle
finishes in seconds.
Tier 2: under independent RNG, first two moments match the closed forms within
5 sigma and a two-sample KS test against oracle CPU draws passes at ALPHA.

Tolerance note (DESIGN.md "Numerical contract"): in the normal-approximation
regime (b > 170) the reference's variance expression pg_m2 - pg_m1^2
(PolyaGamma.cpp:231-239 via LogitWrapper.cpp:143-145) cancels catastrophically as
z -> 0 (tanh(z/2) - z/2); a 1-ulp difference between glibc's and CUDA's tanh is
amplified by ~3/(z/2)^2, so there the bar is REL * max(1, amplification).
"""
import os

import numpy as np
import pytest
from scipy import special, stats

from oracle.loader import make_tape

pytestmark = pytest.mark.gpu

REL = 1e-12
ALPHA = 1e-3
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pg_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def tape_of(g, p):
    return {k: g[f"{p}_t{k}"] for k in "ueng" if f"{p}_t{k}" in g}


def assert_close(a, b, rel=REL, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan), "NaN pattern differs"
    tol = rel * np.abs(b[~nan])
    if scale is not None:
        tol = tol * np.maximum(1.0, scale[~nan])
    bad = np.abs(a[~nan] - b[~nan]) > tol
    assert not bad.any(), (int(bad.sum()), a[~nan][bad][:4], b[~nan][bad][:4])


def normal_regime_amplification(h, z):
    """Condition number of the reference's normal-regime variance w.r.t. tanh."""
    zh = np.abs(z) * 0.5
    amp = np.ones_like(zh)
    sel = (h > 170) & (zh < 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp[sel] = 4.0 * np.tanh(zh[sel]) / np.abs(np.tanh(zh[sel]) - zh[sel])
    amp[sel & (zh == 0)] = 1.0
    return amp


# ----------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------

def test_philox_known_answers(engine):
    assert engine.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == \
        [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert engine.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert engine.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_special_functions(engine):
    x = np.concatenate([np.linspace(-40, 8, 2001), [-1e3, -300.0]])
    np.testing.assert_allclose(engine.specfun("log_p_norm", x), special.log_ndtr(x), rtol=2e-13, atol=1e-16)
    xs = x[x > -37]
    got, want = engine.specfun("p_norm", xs), special.ndtr(xs)
    assert np.all(np.abs(got - want) <= 1e-15 * np.maximum(100, 2 * xs * xs) * want)
    rng = np.random.default_rng(0)
    a = np.concatenate([np.full(500, 0.5), rng.uniform(1, 4, 2000), rng.uniform(13, 170, 4000)])
    xx = a * rng.uniform(0.2, 3.0, a.size)
    got = engine.specfun("p_gamma_rate", xx, a, np.ones_like(a))
    assert np.max(np.abs(got - special.gammainc(a, xx))) < 2e-14
    g = np.linspace(1, 170, 500)
    np.testing.assert_allclose(engine.specfun("lgamma", g), special.gammaln(g), rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(engine.specfun("tgamma", g), special.gamma(g), rtol=1e-13)
    xi, mu, lam = rng.uniform(0.05, 1.1, 500), rng.uniform(0.05, 5, 500), rng.uniform(1, 170, 500)
    np.testing.assert_allclose(engine.specfun("p_igauss", xi, mu, lam),
                               stats.invgauss.cdf(xi, mu / lam, scale=lam), rtol=1e-9, atol=1e-300)


def test_moments_and_v_eval_match_golden(engine, gold):
    m1, m2 = engine.pg_moments(gold["mom_b"], gold["mom_z"])
    assert_close(m1, gold["mom_m1"], 1e-14)
    assert_close(m2, gold["mom_m2"], 1e-13)
    # v is only ever used additively (t = v/2 + z^2/2, InvertY.cpp Newton stops at |dv| <= 1e-9):
    # absolute agreement is the meaningful bar near v = 0
    np.testing.assert_allclose(engine.v_eval(gold["vev_y"]), gold["vev_v"], rtol=1e-12, atol=1e-14)


def test_saddle_point_tables_match_mpmath_golden(engine, oracle):
    """V(y) = y^-1(y) and G(y) = log cos_rt(V(y)) as the saddle-point kernels take them (degree-9
    tables, reference Newton where its iterates clamp) against 40-digit values
    (tests/golden/make_sp_golden.py)."""
    g = dict(np.load(os.path.join(os.path.dirname(GOLD), "sp_tables_golden.npz")))
    y = g["y"]
    vt = engine.specfun("sp_v_table", y)
    on_table = ~np.isnan(vt)
    assert on_table.mean() > 0.98                       # random y: the table path is the path taken
    assert np.max(np.abs(vt[on_table] - g["y_v"][on_table]) / np.maximum(1, np.abs(g["y_v"][on_table]))) < 4e-16
    gg = engine.specfun("sp_log_cos_rt", y)
    assert np.max(np.abs(gg[on_table] - g["y_g"][on_table]) / np.maximum(1, np.abs(g["y_g"][on_table]))) < 4e-16
    # off the table the reference's own iteration decides (its iterates clamp to a 7-digit grid
    # there, so the answer is NOT the root): the engine has to agree with the reference, not mpmath
    off = ~on_table
    v = engine.v_eval(y)
    assert np.array_equal(v[on_table], vt[on_table])
    vo = np.array([oracle.v_eval(float(t)) for t in y[off]])
    np.testing.assert_allclose(v[off], vo, rtol=1e-12, atol=1e-14)
    cos_rt = np.where(vo >= 0, np.cos(np.sqrt(np.abs(vo))), np.cosh(np.sqrt(np.abs(vo))))
    np.testing.assert_allclose(gg[off], np.log(cos_rt), rtol=1e-12, atol=1e-14)
    # hugging the reference's grid points and y = 1 the table path must decline: there the reference's
    # answer is its grid bracket / series branch, not the root
    near = engine.specfun("sp_v_table", g["y_near"])
    d = np.abs(np.log2(g["y_near"]) * 10 - np.round(np.log2(g["y_near"]) * 10))
    assert np.all(np.isnan(near[d < 2e-5]))
    assert np.all(np.isnan(near[np.abs(g["y_near"] - 1) < 5e-7]))


def test_saddle_point_weight_functions(engine):
    """Forward continued fraction of Gamma(a,x) e^x x^-a and the direct inverse-Gaussian CDF."""
    rng = np.random.default_rng(5)
    a = rng.uniform(1, 171, 4000)
    x = np.maximum(a + 1.0, a * rng.uniform(1.0, 2.5, a.size))
    got = engine.specfun("upper_gamma_cf", x, a)
    want = np.exp(np.log(special.gammaincc(a, x)) + special.gammaln(a) + x - a * np.log(x))
    ok = special.gammaincc(a, x) > 1e-280
    np.testing.assert_allclose(got[ok], want[ok], rtol=2e-12)      # scipy's own Q carries ~1e-13
    xi, mu, lam = rng.uniform(0.05, 1.1, 2000), rng.uniform(0.05, 5, 2000), rng.uniform(1, 170, 2000)
    np.testing.assert_allclose(engine.specfun("p_igauss_direct", xi, mu, lam),
                               engine.specfun("p_igauss", xi, mu, lam), rtol=2e-12, atol=1e-300)


def test_saddle_point_pl_estimate(engine):
    """The binned path decides U < pl against an fp32 estimate of pl with a band of 2e-5 and asks
    the fp64 weights only inside the band: the estimate has to sit well inside that band wherever
    one is offered, over the shapes and tilts the regime sees (and beyond)."""
    rng = np.random.default_rng(11)
    n = np.concatenate([rng.uniform(13, 170, 60000), rng.integers(14, 171, 20000).astype(float),
                        rng.uniform(1.5, 13, 5000), rng.uniform(170, 400, 2000)])
    z = np.concatenate([rng.uniform(-5, 5, 70000), rng.uniform(-30, 30, 17000)])
    pl = engine.specfun("sp_pl", n, z)
    est = engine.specfun("sp_pl_estimate", n, z)
    have = ~np.isnan(est)
    core = (n > 13) & (n <= 170) & (np.abs(z) <= 5)
    assert have[core].mean() > 0.999            # the benchmark regime runs on the estimate
    ok = n <= 170       # beyond, the reference's Gamma(n) overflows and its pl is 0 or NaN (reproduced as is)
    assert np.all(np.isfinite(pl[ok])) and np.all((pl[ok] >= 0) & (pl[ok] <= 1))
    assert not have[n > 171].any()
    assert np.max(np.abs(est[have] - pl[have])) < 5e-6


def test_alternate_pr_estimate(engine):
    """Same scheme for the alternate sampler's right-piece mass (chunk shapes 1 .. 4)."""
    rng = np.random.default_rng(12)
    h = np.concatenate([rng.uniform(1, 4, 60000), np.full(20000, 4.0), rng.integers(1, 5, 5000).astype(float)])
    z = np.concatenate([rng.uniform(-5, 5, 70000), rng.uniform(-40, 40, 15000)])
    pr = engine.specfun("alt_pr", h, z)
    est = engine.specfun("alt_pr_estimate", h, z)
    have = ~np.isnan(est)
    assert have[np.abs(z) <= 5].mean() > 0.99
    assert np.all(np.isfinite(pr)) and np.all((pr >= 0) & (pr <= 1))
    assert np.max(np.abs(est[have] - pr[have])) < 5e-6


# ----------------------------------------------------------------------------------
# tier 1: golden vectors made from the reference itself
# ----------------------------------------------------------------------------------

def test_golden_tapes(engine, gold):
    g = gold
    x, tr = engine.rpg_tape("devroye", g["dev_n"], g["dev_z"], tape_of(g, "dev"))
    assert np.array_equal(tr, g["dev_trace"]); assert_close(x, g["dev_x"])
    x, tr = engine.rpg_tape("alt", g["alt_h"], g["alt_z"], tape_of(g, "alt"))
    assert np.array_equal(tr, g["alt_trace"]); assert_close(x, g["alt_x"])
    x, tr, it = engine.rpg_tape("sp", g["sp_h"], g["sp_z"], tape_of(g, "sp"))
    assert np.array_equal(tr, g["sp_trace"]); assert np.array_equal(it, g["sp_iter"])
    assert_close(x, g["sp_x"])
    x, tr = engine.rpg_tape("gamma", g["gam_h"], g["gam_z"], tape_of(g, "gam"), trunc=64)
    assert np.array_equal(tr, g["gam_trace"]); assert_close(x, g["gam_x"])
    x, tr = engine.rpg_tape("hybrid", g["hyb_h"], g["hyb_z"], tape_of(g, "hyb"))
    assert np.array_equal(tr, g["hyb_trace"])
    assert_close(x, g["hyb_x"], scale=normal_regime_amplification(g["hyb_h"], g["hyb_z"]))


def test_golden_philox_streams(engine, gold):
    g = gold
    assert_close(engine.rpg_seeded("devroye", g["pdev_n"], g["pdev_z"], 20240001, call_id=3, obs0=7), g["pdev_x"])
    assert_close(engine.rpg_seeded("alt", g["palt_h"], g["palt_z"], 20240002, obs0=1 << 33), g["palt_x"])
    x, it = engine.rpg_seeded("sp", g["psp_h"], g["psp_z"], 20240003)
    assert_close(x, g["psp_x"]); assert np.array_equal(it, g["psp_iter"])
    assert_close(engine.rpg_seeded("gamma", g["pgam_h"], g["pgam_z"], 20240004, trunc=200), g["pgam_x"])
    assert_close(engine.rpg_seeded("hybrid", g["phyb_h"], g["phyb_z"], 20240005, call_id=9), g["phyb_x"],
                 scale=normal_regime_amplification(g["phyb_h"], g["phyb_z"]))


# ----------------------------------------------------------------------------------
# tier 1 at scale: >= 1e6 (z, tape) cases for PG(1,z), z in [-50,50] U {0}
# ----------------------------------------------------------------------------------

def test_devroye_tape_parity_one_million(engine, oracle):
    rng = np.random.default_rng(20240001)
    total = mismatched = dry = 0
    for chunk in range(4):
        m = 250_000
        z = rng.uniform(-50, 50, m)
        z[:50_000] = rng.uniform(-5, 5, 50_000)
        z[50_000:50_100] = 0.0
        n = np.ones(m, dtype=np.int32)
        n[-10_000:] = rng.integers(0, 4, 10_000)
        tape = make_tape(m, lu=24, le=24, ln=8, seed=100 + chunk)
        xa, ta = engine.rpg_tape("devroye", n, z, tape)
        xb, tb = oracle.rpg_devroye(n, z, tape=tape, trace=True, nthreads=8)
        mismatched += int((ta != tb).any(axis=1).sum())
        dry += int(tb[:, 4].sum())
        assert_close(xa, xb)
        total += m
    assert total >= 1_000_000
    assert mismatched == 0
    assert dry < total // 100


@pytest.mark.parametrize("method", ["alt", "sp", "hybrid", "gamma"])
def test_tape_parity_other_regimes(engine, oracle, method):
    rng = np.random.default_rng({"alt": 1, "sp": 2, "hybrid": 3, "gamma": 4}[method])
    m = 100_000 if method != "gamma" else 4000
    z = rng.uniform(-12, 12, m)
    z[:64] = 0.0
    if method == "alt":
        h = rng.uniform(1, 30, m); h[:1000] = rng.integers(1, 31, 1000)
        tape = make_tape(m, lu=96, le=96, ln=24, seed=21)
    elif method == "sp":
        h = rng.uniform(1, 170, m); h[:1000] = rng.integers(14, 171, 1000)
        tape = make_tape(m, lu=24, le=24, ln=8, seed=22)
    elif method == "gamma":
        h = rng.uniform(0.05, 4, m)
        tape = make_tape(m, lg=100, g_shape=h, seed=24)
    else:
        h = np.where(rng.random(m) < 0.5, rng.uniform(1, 200, m), rng.integers(1, 201, m).astype(float))
        tape = make_tape(m, lu=96, le=96, ln=24, seed=23)
    kw = {"trunc": 100} if method == "gamma" else {}
    got = engine.rpg_tape(method, h, z, tape, **kw)
    want = getattr(oracle, "rpg_" + method)(h, z, tape=tape, trace=True, nthreads=8, **kw)
    assert np.array_equal(got[1], want[1]), int((got[1] != want[1]).any(axis=1).sum())
    scale = normal_regime_amplification(h, z) if method == "hybrid" else None
    assert_close(got[0], want[0], scale=scale)
    if method == "sp":
        assert np.array_equal(got[2], want[2])
    assert want[1][:, 4].mean() < 0.02


# ----------------------------------------------------------------------------------
# stream contract at scale: same Philox streams on GPU and in the oracle
# ----------------------------------------------------------------------------------

@pytest.mark.parametrize("method,num", [("devroye", 2_000_000), ("alt", 1_000_000), ("sp", 1_000_000),
                                        ("hybrid", 1_000_000), ("gamma", 20_000)])
def test_philox_stream_parity(engine, oracle, method, num):
    rng = np.random.default_rng(7)
    z = rng.uniform(-5, 5, num)
    if method == "devroye":
        shape = rng.integers(0, 5, num).astype(np.int32)
    elif method == "alt":
        shape = rng.uniform(1, 13, num)
    elif method == "sp":
        shape = rng.uniform(13, 170, num)
    elif method == "gamma":
        shape = rng.uniform(0.05, 3, num)
    else:
        shape = np.where(rng.random(num) < 0.5, rng.uniform(0.5, 200, num),
                         rng.integers(1, 201, num).astype(float))
        shape[:1000] = rng.uniform(0.01, 1.0, 1000)
    obs0 = (1 << 32) - 1000   # straddle the 32-bit boundary of the counter
    got = engine.rpg_seeded(method, shape, z, seed=0xC0FFEE, call_id=5, obs0=obs0)
    want = getattr(oracle, "rpg_" + method)(shape, z, seed=0xC0FFEE, call_id=5, obs0=obs0, nthreads=8)
    if method == "sp":
        assert np.array_equal(got[1], want[1])
        got, want = got[0], want[0]
    scale = normal_regime_amplification(shape, z) if method == "hybrid" else None
    assert_close(got, want, scale=scale)


def test_philox_stream_parity_binned_path_wide_tilt(engine, oracle):
    """The regime-binned rpg_hybrid kernels over tilts far outside the benchmark's range: there the
    fp32 mass estimates are not offered, the gamma tail leaves the continued-fraction form, the
    saddle-point abscissae leave the V/G tables and Gamma(n) overflows -- every hand-over to the
    reference-faithful forms is exercised, draw for draw against the oracle."""
    rng = np.random.default_rng(17)
    num = 400_000
    z = np.concatenate([rng.uniform(-40, 40, num // 2), rng.normal(0, 1e-3, num // 4), rng.uniform(-90, 90, num // 4)])
    z[:100] = 0.0
    shape = np.where(rng.random(num) < 0.5, rng.uniform(0.5, 200, num), rng.integers(1, 201, num).astype(float))
    got = engine.rpg_seeded("hybrid", shape, z, seed=0xBEEF, call_id=1, obs0=123)
    want = oracle.rpg_hybrid(shape, z, seed=0xBEEF, call_id=1, obs0=123, nthreads=8)
    assert_close(got, want, scale=normal_regime_amplification(shape, z))


def test_binned_path_is_deterministic(engine):
    """The loop kernels hand draws from thread to thread through shared-memory slots and sort them
    every trip; which thread advances which draw depends on scheduling.  Results must not: three
    runs of the same 2M-draw batch are bit-identical (compute-sanitizer is not available on the
    pool, this is the race check that is)."""
    rng = np.random.default_rng(23)
    num = 2_000_000
    z = rng.uniform(-6, 6, num)
    h = np.where(rng.random(num) < 0.5, rng.uniform(0.5, 200, num), rng.integers(1, 201, num).astype(float))
    a = engine.rpg_seeded("hybrid", h, z, seed=5, call_id=3)
    for _ in range(2):
        assert np.array_equal(a, engine.rpg_seeded("hybrid", h, z, seed=5, call_id=3))
    assert np.all(np.isfinite(a)) and np.all(a > 0)


def test_results_independent_of_chunking_and_sharding(engine):
    """obs0 keys the stream by global observation index: two half batches equal one full batch."""
    rng = np.random.default_rng(3)
    num = 300_000
    z = rng.uniform(-5, 5, num)
    h = rng.uniform(0.5, 200, num)
    full = engine.rpg_seeded("hybrid", h, z, seed=99, call_id=1, obs0=10)
    half = num // 2
    a = engine.rpg_seeded("hybrid", h[:half], z[:half], seed=99, call_id=1, obs0=10)
    b = engine.rpg_seeded("hybrid", h[half:], z[half:], seed=99, call_id=1, obs0=10 + half)
    assert np.array_equal(full, np.concatenate([a, b]))


# ----------------------------------------------------------------------------------
# tier 2: independent RNG -- moments and KS against oracle CPU draws
# ----------------------------------------------------------------------------------

@pytest.mark.parametrize("b", [1.0, 2.0, 3.0, 1.5, 3.7, 10.0, 14.0, 50.5, 170.0, 250.0, 0.5])
def test_moments_and_ks_vs_oracle(engine, oracle, b):
    n = 400_000 if b >= 1 else 20_000
    for z in (0.0, 1.3, 4.0):
        zz = np.full(n, z)
        hh = np.full(n, b)
        if b == 3.0:   # sum-of-PG(1) path through rpg_devroye
            x = engine.rpg_seeded("devroye", hh.astype(np.int32), zz, seed=1234 + int(10 * z))
            y = oracle.rpg_devroye(hh.astype(np.int32)[: n // 4], zz[: n // 4], seed=977, nthreads=8)
        else:
            x = engine.rpg_seeded("hybrid", hh, zz, seed=1234 + int(10 * z))
            y = oracle.rpg_hybrid(hh[: n // 4], zz[: n // 4], seed=977, nthreads=8)
        if z == 0:
            m, v = b / 4.0, b / 24.0
        else:
            m = b / (2 * z) * np.tanh(z / 2)
            v = b / (4 * z ** 3) * (np.sinh(z) - z) / np.cosh(z / 2) ** 2
        if b < 1:   # the 200-term truncated sum of gammas is biased low by construction
            m, v = y.mean(), y.var()
            assert abs(x.mean() - m) < 5 * np.sqrt(v / n + v / len(y))
        else:
            assert abs(x.mean() - m) < 5 * np.sqrt(v / n)
            assert abs(x.var() - v) < 5 * np.sqrt(11 * v * v / n)
        assert stats.ks_2samp(x, y).pvalue > ALPHA


def test_ks_by_z_bucket_pg1(engine, oracle):
    rng = np.random.default_rng(5)
    n = 1_000_000
    z = rng.uniform(-5, 5, n)
    one = np.ones(n, dtype=np.int32)
    x = engine.rpg_seeded("devroye", one, z, seed=31337)
    y = oracle.rpg_devroye(one, z, seed=42, nthreads=8)
    edges = np.linspace(0, 5, 11)
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (np.abs(z) >= lo) & (np.abs(z) < hi)
        assert stats.ks_2samp(x[sel], y[sel]).pvalue > ALPHA


# ----------------------------------------------------------------------------------
# edge cases and the drop-in (R-facing) entry points
# ----------------------------------------------------------------------------------

def test_edge_cases(engine, oracle):
    # empty batch
    assert engine.rpg_devroye(0, 1, 0.0).size == 0
    # n == 0 / h == 0 / h <= 0 give exactly 0 (LogitWrapper.cpp:76-79, 159-161)
    z = np.array([0.0, 1.0, -2.0, 700.0, -1e3, 1e-300])
    assert np.all(engine.rpg_seeded("devroye", np.zeros(6, np.int32), z, 1) == 0)
    assert np.all(engine.rpg_seeded("hybrid", np.array([0, -1.0, -0.0, 0, 0, 0]), z, 1) == 0)
    # extreme tilts stay finite, positive and equal to the oracle
    h = np.array([1.0, 2.0, 3.5, 20.0, 1.0, 1.0])
    got = engine.rpg_seeded("hybrid", h, z, seed=5)
    want = oracle.rpg_hybrid(h, z, seed=5)
    assert np.all(np.isfinite(got)) and np.all(got > 0)
    assert_close(got, want)
    # rpg_sp leaves iter untouched where h == 0 (LogitWrapper.cpp:119-122)
    from bayeslogit_b200 import _lib
    import ctypes as C
    hh = np.array([20.0, 0.0, 30.0]); zz = np.zeros(3); x = np.full(3, -7.0)
    it = np.array([-5, -5, -5], dtype=np.int32)
    _lib.lib().rpg_sp(x.ctypes.data, hh.ctypes.data, zz.ctypes.data, C.byref(C.c_int(3)), it.ctypes.data)
    _lib.check()
    assert it[1] == -5 and it[0] >= 1 and it[2] >= 1 and x[1] == 0.0
    # Alt sampler: h < 1 -> 0 (PolyaGammaAlt.cpp:207-210)
    assert engine.rpg_seeded("alt", np.array([0.5]), np.array([1.0]), 1)[0] == 0.0
    # n < 1 is clamped to 1 under NTHROW (PolyaGamma.cpp:128-135)
    a = engine.rpg_seeded("devroye", np.array([-3], np.int32), np.array([0.7]), 9)
    b = engine.rpg_seeded("devroye", np.array([1], np.int32), np.array([0.7]), 9)
    assert a[0] == b[0]


def test_tape_running_dry_is_flagged(engine, oracle):
    m = 5000
    rng = np.random.default_rng(8)
    z = rng.uniform(-5, 5, m)
    tape = make_tape(m, lu=6, le=6, ln=3, seed=12)
    n = np.ones(m, dtype=np.int32)
    xa, ta = engine.rpg_tape("devroye", n, z, tape)
    xb, tb = oracle.rpg_devroye(n, z, tape=tape, trace=True)
    assert np.array_equal(np.isnan(xa), tb[:, 4] == 1)
    ok = tb[:, 4] == 0
    assert np.array_equal(ta[ok], tb[ok])
    assert_close(xa[ok], xb[ok])
    assert 0 < (~ok).sum() < m


def test_dropin_entry_points_mirror_r_wrappers(engine, oracle, capsys):
    engine.set_seed(2024)
    a = engine.rpg_devroye(1000, 1, 0.5)
    b = engine.rpg_devroye(1000, 1, 0.5)
    assert not np.array_equal(a, b)            # the call counter advances like an RNG state
    engine.set_seed(2024)
    assert np.array_equal(a, engine.rpg_devroye(1000, 1, 0.5))   # set.seed analogue
    # the drop-in call is the stream (seed, call 0, obs 0..): same numbers as the oracle
    want = oracle.rpg_devroye(np.ones(1000, np.int32), np.full(1000, 0.5), seed=2024, call_id=0)
    assert_close(a, want)
    # recycling of h and z to length num, as R's array(h, num)
    engine.set_seed(7)
    x = engine.rpg(6, [1.0, 20.0], [0.0, 1.0, 2.0])
    want = oracle.rpg_hybrid(np.array([1.0, 20.0] * 3), np.array([0.0, 1.0, 2.0] * 2), seed=7)
    assert_close(x, want)
    # validation messages and NA returns of LogitWrapper.R:15-22,39-42,59-62,107-110
    assert np.isnan(engine.rpg(3, -1.0, 0.0)); assert "h must be > 0." in capsys.readouterr().out
    assert np.isnan(engine.rpg_alt(3, 0.5, 0.0)); assert "h must be >= 1." in capsys.readouterr().out
    assert np.isnan(engine.rpg_devroye(3, -1, 0.0))
    assert np.isnan(engine.rpg_gamma(3, 1.0, 0.0, trunc=0))
    out = engine.rpg_sp(50, 30.0, 1.0, track_iter=True)
    assert out["samp"].shape == (50,) and np.all(out["iter"] >= 1)
    g = engine.rpg_gamma(2000, 2.0, 1.0, trunc=200)
    assert abs(g.mean() - 2.0 / 2 * np.tanh(0.5)) < 0.05


@pytest.mark.parametrize("method", ["devroye", "hybrid"])
def test_large_batch_chunk_pipeline(engine, oracle, method):
    """Many pipeline chunks (ramped schedule K/8, K/4, K/2, K ..., K/2, K/4, K/8 with K = 4M; three slots in
    rotation) through the host-pointer ABI: draws on both sides of every chunk boundary equal the oracle's
    for the same global observation index, and the whole batch equals the device-resident entry point's."""
    import torch
    from bayeslogit_b200 import _lib
    M = 1 << 20
    num = 34 * M + 12345
    rng = np.random.default_rng(1)
    z = rng.uniform(-5, 5, num)
    if method == "devroye":
        shape = np.ones(num, dtype=np.int32)
    else:
        shape = np.where(rng.random(num) < 0.5, rng.uniform(0.5, 200, num), rng.integers(1, 201, num).astype(float))
    x = engine.rpg_seeded(method, shape, z, seed=77, call_id=2)
    fn = getattr(oracle, "rpg_" + method)
    scale_all = normal_regime_amplification(shape, z) if method == "hybrid" else None
    import ctypes as C
    buf = (C.c_int64 * 256)()
    k = _lib.lib().bl_probe_pipeline_schedule(num, C.cast(buf, C.c_void_p), 256)
    sizes = list(buf[:k])             # the chunk sizes run_host uses (ramp K/8, K/4, K/2 at both ends, K = 4M)
    assert sum(sizes) == num and len(sizes) >= 10
    bounds = np.cumsum(sizes)[:-1].tolist()
    for i0 in [0] + [b - 2500 for b in bounds] + [num - 5000]:
        want = fn(shape[i0:i0 + 5000], z[i0:i0 + 5000], seed=77, call_id=2, obs0=i0)
        assert_close(x[i0:i0 + 5000], want, scale=None if scale_all is None else scale_all[i0:i0 + 5000])
    dev = torch.device("cuda", 0)
    sd, zd = torch.from_numpy(shape).to(dev), torch.from_numpy(z).to(dev)
    xd = torch.empty(num, dtype=torch.float64, device=dev)
    f_dev = getattr(_lib.lib(), "bl_rpg_%s_dev" % method)
    _lib.check(f_dev(xd.data_ptr(), sd.data_ptr(), zd.data_ptr(), num, 77, 2, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(x, xd.cpu().numpy()), "host pipeline and device-resident batch differ"
    if method == "devroye":
        m = 0.5 / z * np.tanh(z / 2)
        assert abs((x - m).mean()) < 5 * 0.2 / np.sqrt(num)


def test_r_seed_shim(engine):
    """bl_set_seed_r(int*) -- what an R front end calls next to set.seed(): same state as bl_set_seed."""
    import ctypes as C
    from bayeslogit_b200 import _lib
    L = _lib.lib()
    z = np.linspace(-3, 3, 1000)
    seed = C.c_int(4711)
    L.bl_set_seed_r(C.cast(C.byref(seed), C.c_void_p))
    assert L.bl_get_seed() == 4711 and L.bl_get_call_counter() == 0
    a = engine.rpg_devroye(1000, 1, z)
    engine.set_seed(4711)
    b = engine.rpg_devroye(1000, 1, z)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("num", [40_000, 1_500_003])
def test_devroye_regroup_equals_refill(engine, oracle, num):
    """k_devroye_regroup (draws in shared-memory slots, regrouped by proposal piece across the CTA every trip;
    BL_DEVROYE_REGROUP=1, batches >= 2^15) against k_devroye_refill (a lane owns a draw; the default) and the oracle: the same
    stream per observation consumed in the same order, so the same bits -- n_i in {0, 1, 2, 5, -3} (zeros, sums of
    PG(1), the NTHROW clamp) and |z| up to 40 (every proposal piece, both series branches)."""
    import os
    rng = np.random.default_rng(17)
    z = np.where(rng.random(num) < 0.9, rng.uniform(-5, 5, num), rng.uniform(-40, 40, num))
    n = rng.choice(np.array([0, 1, 1, 1, 1, 2, 5, -3], dtype=np.int32), num)
    a = engine.rpg_seeded("devroye", n, z, seed=4242, call_id=3, obs0=17)
    os.environ["BL_DEVROYE_REGROUP"] = "1"
    try:
        b = engine.rpg_seeded("devroye", n, z, seed=4242, call_id=3, obs0=17)
    finally:
        os.environ.pop("BL_DEVROYE_REGROUP", None)
    assert np.array_equal(a, b)
    m = min(num, 200_000)
    want = oracle.rpg_devroye(n[:m], z[:m], seed=4242, call_id=3, obs0=17, nthreads=8)
    assert_close(a[:m], want)
    assert np.all(a[n == 0] == 0.0) and np.all(a[n != 0] > 0)
