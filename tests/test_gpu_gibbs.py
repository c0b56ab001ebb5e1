"""GPU parity tests of the Gibbs sweeps against the CPU
restatements in oracle/gibbs_oracle.c, through the C ABI.

Engine and oracle share the stream contract (iteration t: omega_i from Philox
(seed, obs i, call t), beta from (seed, obs 2^64-1, call t)), so for the same seed the
two CHAINS must agree -- not just their distributions.  Tolerance CHAIN_REL = 1e-8 on beta
and omega over tens of iterations: the per-draw bar is 1e-12, the Gram is an N-term sum
evaluated in a different order on the device (relative 1e-13..1e-12), and the chain map
amplifies that mildly from one iteration to the next.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import loader

pytestmark = pytest.mark.gpu

CHAIN_REL = 1e-8


@pytest.fixture(scope="module")
def gapi(engine):
    from bayeslogit_b200 import gibbs_api
    return gibbs_api


def synth_logit(N, P, seed, binomial=False):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    bt = np.r_[np.abs(rng.normal(0, 0.4, P - 1)), -0.5]
    p = 1 / (1 + np.exp(-X @ bt))
    if binomial:
        n = rng.integers(1, 6, N).astype(float)
        y = rng.binomial(n.astype(int), p) / n
    else:
        n = np.ones(N)
        y = (rng.random(N) < p).astype(float)
    return X, y, n, bt


def close(a, b, rel=CHAIN_REL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    assert np.all(np.isfinite(a)) and err.max() <= rel, float(err.max())


@pytest.mark.parametrize("constrained", [False, True])
@pytest.mark.parametrize("N,P,binomial", [(3000, 7, False), (1500, 64, False), (2000, 5, True), (700, 70, False), (1501, 32, False), (3000, 128, False),
                                          (40000, 72, False), (4301, 33, False), (7300, 63, False), (1200, 2, False), (1100, 1, False)])
def test_logit_chain_matches_oracle(gapi, constrained, N, P, binomial):
    # N = 40000, P = 72: enough rows for the cost-proportional slab counts of the P > 64 Gram (diagonal
    # tiles cut into fewer slabs than off-diagonal ones; smaller N caps both at N / 128).
    # P = 128 runs the beta draw out of global scratch (2 P^2 doubles exceed shared memory) and the
    # four-way column rotation of k_xtv_stream.  N is kept >= ~20 P: the constrained draw starts on
    # the constraint boundary (beta = 0), where its truncation windows are a few ulps wide and the
    # chain map amplifies last-bit differences of erfc without bound when N is only a few P.
    # P = 33 / 63: the second register of a lane's pair is only partly populated (identity-padded fast beta draws);
    # P = 1, 2: no / one constrained coefficient.
    X, y, n, _ = synth_logit(N, P, 10 + P, binomial)
    m0 = np.linspace(-0.1, 0.1, P)
    P0 = 0.5 * np.eye(P) + 0.01
    samp, burn = (12, 6) if P < 32 else (5, 3)
    flags = 0 if constrained else gapi.PLAIN_BETA
    w, b = gapi.logit_gibbs(y, X, n, m0, P0, samp, burn, seed=77, flags=flags)
    wo, bo = loader.logit_gibbs(y, X, n, m0, P0, samp, burn, seed=77, constrained=constrained)
    close(b, bo)
    close(w, wo)
    if constrained:
        assert np.all(b[:, :-1] >= 0)


@pytest.mark.parametrize("constrained", [False, True])
@pytest.mark.parametrize("N,P", [(1500, 64), (20_001, 40), (1_000_000, 64)])
def test_one_pass_sweep_matches_oracle(gapi, constrained, N, P):
    """The one-pass sweep kernel against the oracle directly (N = 1 000 000, P = 64 is BASELINE config 3's shape)."""
    X, y, n, _ = synth_logit(N, P, 31 + P)
    m0 = np.zeros(P)
    P0 = 0.05 * np.eye(P)
    samp, burn = (2, 1) if N > 100_000 else (5, 3)
    flags = gapi.ONE_PASS | (0 if constrained else gapi.PLAIN_BETA)
    w, b = gapi.logit_gibbs(y, X, n, m0, P0, samp, burn, seed=91, flags=flags)
    wo, bo = loader.logit_gibbs(y, X, n, m0, P0, samp, burn, seed=91, constrained=constrained)
    close(b, bo); close(w, wo)


def test_logit_burn_zero_and_no_w(gapi):
    X, y, n, _ = synth_logit(1000, 4, 5)
    P0 = np.eye(4)
    w, b = gapi.logit_gibbs(y, X, n, np.zeros(4), P0, 6, 0, seed=3, flags=gapi.PLAIN_BETA)
    wo, bo = loader.logit_gibbs(y, X, n, np.zeros(4), P0, 6, 0, seed=3, constrained=False)
    close(b, bo); close(w, wo)
    w2, b2 = gapi.logit_gibbs(y, X, n, np.zeros(4), P0, 6, 0, seed=3, flags=gapi.PLAIN_BETA, keep_w=False)
    assert w2 is None and np.array_equal(b2, b)


@pytest.mark.parametrize("constrained", [False, True])
@pytest.mark.parametrize("chains,N,P", [(7, 900, 5), (3, 2101, 32), (2, 2115, 70)])
def test_batched_chains_match_oracle_chain_by_chain(gapi, constrained, chains, N, P):
    """BASELINE config 5 (independent chains) at oracle-sized shapes: chain c
    of the batch equals the oracle's chain on its rows with seed + c, and the engine's own
    single-chain entry point with that seed.  N not a multiple of 32 and P > 64 (several Gram
    tiles) are covered.  (N >= 30 P: with N = 7 P the constrained chain's map amplifies last-bit differences
    from 1e-16 to 1e-6 within eight iterations -- conditioning of the test problem, see the note at
    test_logit_chain_matches_oracle.)"""
    data = [synth_logit(N, P, 100 + 7 * c + P, binomial=(c % 2 == 1)) for c in range(chains)]
    X = np.stack([d[0] for d in data]); y = np.stack([d[1] for d in data]); n = np.stack([d[2] for d in data])
    m0 = np.linspace(-0.1, 0.1, P)
    P0 = 0.5 * np.eye(P) + 0.01
    samp, burn = (8, 4) if P < 64 else (4, 2)
    flags = 0 if constrained else gapi.PLAIN_BETA
    b = gapi.logit_chains(y, X, n, m0, P0, samp, burn, seed=500, flags=flags)
    assert b.shape == (chains, samp, P)
    for c in range(chains):
        _, bo = loader.logit_gibbs(y[c], X[c], n[c], m0, P0, samp, burn, seed=500 + c, constrained=constrained)
        close(b[c], bo)
        _, b1 = gapi.logit_gibbs(y[c], X[c], n[c], m0, P0, samp, burn, seed=500 + c, flags=flags, keep_w=False)
        close(b[c], b1)
    assert not np.allclose(b[0], b[1][:, :P])          # different data and streams: different chains


def test_mlogit_chain_matches_oracle(gapi):
    rng = np.random.default_rng(4)
    N, P, J = 2500, 6, 4
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    B = rng.normal(0, 0.7, (P, J - 1))
    eta = np.c_[X @ B, np.zeros(N)]
    pr = np.exp(eta); pr /= pr.sum(1, keepdims=True)
    cat = (pr.cumsum(1) < rng.random(N)[:, None]).sum(1)
    Y = np.eye(J)[cat][:, :J - 1]
    m0 = rng.normal(0, 0.1, (P, J - 1))
    P0 = np.stack([(0.3 + 0.1 * j) * np.eye(P) for j in range(J - 1)], axis=2)
    w, b = gapi.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 6, 3, seed=21)
    wo, bo = loader.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 6, 3, seed=21)
    close(b, bo); close(w, wo)


def synth_mlogit(N, P, J, seed, scale=0.7):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    B = rng.normal(0, scale, (P, J - 1))
    eta = np.c_[X @ B, np.zeros(N)]
    pr = np.exp(eta); pr /= pr.sum(1, keepdims=True)
    cat = (pr.cumsum(1) < rng.random(N)[:, None]).sum(1)
    Y = np.eye(J)[cat][:, :J - 1]
    return X, Y, B


def test_mlogit_chain_matches_oracle_at_the_benchmarked_shape(gapi):
    """BASELINE config 5a's kernel instantiations (J = 10, P = 32) at an oracle-sized N: X'(Omega c_j) through
    k_xtv_stream<4, 1> (P a power of two, weight mode omega * c), the packed P = 32 Gram, the mvn beta
    draw at P = 32, nine category updates per iteration (MultLogit.hpp:242-258, 293-307)."""
    N, P, J = 20_011, 32, 10
    X, Y, _ = synth_mlogit(N, P, J, 41, scale=0.25)
    rng = np.random.default_rng(5)
    m0 = rng.normal(0, 0.05, (P, J - 1))
    P0 = np.stack([(0.5 + 0.05 * j) * np.eye(P) for j in range(J - 1)], axis=2)
    w, b = gapi.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 3, 2, seed=77)
    wo, bo = loader.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 3, 2, seed=77)
    close(b, bo); close(w, wo)


@pytest.mark.parametrize("N,P,active", [(20_000, 64, False), (20_000, 64, True), (8_000, 32, True), (5_000, 7, True)])
def test_constrained_draw_speculation_is_bit_identical(gapi, N, P, active):
    """cta_constrained_sweeps_spec (whole CTA, speculating that every coordinate accepts its first rejection normal,
    replaying the chain of FMAs in the sequential order) against the one-warp sequential sweeps (BL_BETA_NO_SPEC):
    same variates, same operation order, same bits (Logit.hpp:366-399).  `active`: a model whose constraints bind
    (negative coefficients), so misses, re-speculation and the sequential fallback all run."""
    import os
    rng = np.random.default_rng(900 + P + int(active))
    X = np.c_[rng.standard_normal((N, P - 1)) / np.sqrt(P), np.ones(N)]
    bt = rng.normal(0, 1.0, P) if active else np.abs(rng.normal(0, 1.0, P)) + 0.5
    y = (rng.random(N) < 1 / (1 + np.exp(-X @ bt))).astype(float)
    m0, P0 = np.zeros(P), 0.01 * np.eye(P)
    w1, b1 = gapi.logit_gibbs(y, X, np.ones(N), m0, P0, 5, 2, seed=31, keep_w=False)
    os.environ["BL_BETA_NO_SPEC"] = "1"
    try:
        w2, b2 = gapi.logit_gibbs(y, X, np.ones(N), m0, P0, 5, 2, seed=31, keep_w=False)
    finally:
        del os.environ["BL_BETA_NO_SPEC"]
    assert np.array_equal(b1, b2)


def test_mlogit_fused_psi_and_next_offsets_equal_the_separate_kernels(gapi):
    """k_xbeta_mma<true> (psi_j, exp(psi_j) cached, offsets and tilt of the next category from one pass over X)
    against the psi kernel + k_mlogit_offsets per category (BL_MLOGIT_UNFUSED): same exp() of the same numbers in
    the same order (MultLogit.hpp:246-258), so the chains agree bit for bit."""
    import os
    N, P, J = 6_007, 32, 5
    X, Y, _ = synth_mlogit(N, P, J, 43, scale=0.3)
    rng = np.random.default_rng(6)
    m0 = rng.normal(0, 0.05, (P, J - 1))
    P0 = np.stack([(0.5 + 0.05 * j) * np.eye(P) for j in range(J - 1)], axis=2)
    w1, b1 = gapi.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 4, 2, seed=79)
    os.environ["BL_MLOGIT_UNFUSED"] = "1"
    try:
        w2, b2 = gapi.mlogit_gibbs(Y, X, np.ones(N), m0, P0, 4, 2, seed=79)
    finally:
        del os.environ["BL_MLOGIT_UNFUSED"]
    assert np.array_equal(b1, b2) and np.array_equal(w1, w2)


@pytest.mark.parametrize("N,P", [(30_000, 32), (30_011, 64), (30_000, 256)])
def test_nb_chain_matches_oracle_at_wide_P(gapi, N, P):
    """BASELINE config 4's kernel instantiations at an oracle-sized N: k_xtv_stream<log2(P/2), 2> (weights
    kappa + omega log d; P = 256 rotates four column pairs per lane), the P > 64 multi-tile Gram under
    hybrid-sampler weights, the plain beta draw out of global scratch (P = 256), the regime-binned
    rpg_hybrid on b = y + d (NBPG-logmean.R:13-34)."""
    rng = np.random.default_rng(60 + P)
    d = 10.0
    X = np.c_[rng.standard_normal((N, P - 1)) / np.sqrt(P), np.ones(N)]
    bt = np.r_[rng.normal(0, 1.0, P - 1), np.log(100.0)]                 # mean count ~ 100: b = y + d in the saddle-point range, a tail > 170
    mu = np.exp(X @ bt)
    y = rng.negative_binomial(d, d / (mu + d)).astype(float)
    assert (y + d > 170).any() and ((y + d > 13) & (y + d <= 170)).mean() > 0.5
    samp = 3
    w, b = gapi.nb_gibbs(y, X, d, np.zeros(P), 0.01 * np.eye(P), samp, seed=18)
    wo, bo = loader.nb_gibbs(y, X, d, np.zeros(P), 0.01 * np.eye(P), samp, seed=18)
    close(b, bo, 1e-7); close(w, wo, 1e-7)


@pytest.mark.parametrize("constrained", [False, True])
def test_logit_chain_matches_oracle_at_full_size(gapi, constrained):
    """BASELINE config 3 at its stated size, N = 1 000 000, P = 64: the multi-wave grids of the fused
    psi + omega pass, the slab counts of the P = 64 Gram and both beta draws against the oracle
    (OpenMP on the host cores: seconds)."""
    N, P = 1_000_000, 64
    X, y, n, _ = synth_logit(N, P, 20240003)
    m0 = np.zeros(P)
    P0 = 0.01 * np.eye(P)
    flags = 0 if constrained else gapi.PLAIN_BETA
    w, b = gapi.logit_gibbs(y, X, n, m0, P0, 2, 1, seed=20240003, flags=flags)
    wo, bo = loader.logit_gibbs(y, X, n, m0, P0, 2, 1, seed=20240003, constrained=constrained)
    close(b, bo); close(w, wo)


def test_batched_chains_match_oracle_at_the_benchmarked_shape(gapi):
    """BASELINE config 5b's per-chain shape (N = 10 000, P = 32) with the reference's constrained draw."""
    chains, N, P = 3, 10_000, 32
    data = [synth_logit(N, P, 300 + c) for c in range(chains)]
    X = np.stack([d[0] for d in data]); y = np.stack([d[1] for d in data]); n = np.stack([d[2] for d in data])
    m0 = np.zeros(P)
    P0 = 0.01 * np.eye(P)
    b = gapi.logit_chains(y, X, n, m0, P0, 4, 2, seed=20240006, flags=0)
    for c in range(chains):
        _, bo = loader.logit_gibbs(y[c], X[c], n[c], m0, P0, 4, 2, seed=20240006 + c, constrained=True)
        close(b[c], bo)


@pytest.mark.parametrize("P", [5, 8])
def test_nb_chain_matches_oracle(gapi, P):
    rng = np.random.default_rng(6)
    N, d = 2000, 3.0
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    bt = np.r_[[0.9, -0.7, 0.6, 0.5], np.zeros(P - 5), 2.6]          # counts 0..500: b = y + d spans Alt/SP/normal
    mu = np.exp(X @ bt)
    y = rng.negative_binomial(d, d / (mu + d)).astype(float)
    assert (y + d > 170).any() and (y + d < 13).any()
    w, b = gapi.nb_gibbs(y, X, d, np.zeros(P), 0.01 * np.eye(P), 6, seed=8)
    wo, bo = loader.nb_gibbs(y, X, d, np.zeros(P), 0.01 * np.eye(P), 6, seed=8)
    close(b, bo, 1e-7); close(w, wo, 1e-7)


def test_nb_chain_with_dispersion_update_matches_oracle(gapi):
    """The full NB.PG.gibbs (dispersion sampled on the device by draw.df): d equal step for step,
    beta and omega within the chain tolerance."""
    rng = np.random.default_rng(16)
    N, P, d = 2500, 4, 4.0
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    bt = np.array([0.8, -0.6, 0.5, 2.2])
    mu = np.exp(X @ bt)
    y = rng.negative_binomial(d, d / (mu + d)).astype(float)
    w, b, ds = gapi.nb_gibbs_df(y, X, np.zeros(P), 0.01 * np.eye(P), 14, 6, seed=8)
    wo, bo, dso = loader.nb_gibbs_df(y, X, np.zeros(P), 0.01 * np.eye(P), 14, 6, seed=8)
    assert np.array_equal(ds, dso) and len(set(ds)) > 1     # it moves, and moves identically
    close(b, bo, 1e-7); close(w, wo, 1e-7)
    # a batch large enough for the regime-binned draw path
    N2 = 40_000
    X2 = np.c_[rng.standard_normal((N2, P - 1)), np.ones(N2)]
    y2 = rng.negative_binomial(d, d / (np.exp(X2 @ bt) + d)).astype(float)
    w, b, ds = gapi.nb_gibbs_df(y2, X2, np.zeros(P), 0.01 * np.eye(P), 5, 3, seed=9, d0=3.0)
    wo, bo, dso = loader.nb_gibbs_df(y2, X2, np.zeros(P), 0.01 * np.eye(P), 5, 3, seed=9, d0=3.0)
    assert np.array_equal(ds, dso)
    close(b, bo, 1e-7); close(w, wo, 1e-7)


def test_nb_chain_with_real_valued_dispersion_matches_oracle(gapi):
    """draw.df.real.mean (NB-Shape.R:86-96): the dispersion walks on the reals, so b = y + d is non-integer and every
    omega comes from the alternate / saddle-point / normal samplers.  d equal step for step (decisions lu < lalpha
    on N-term sums: equal unless a tie to rounding), beta and omega within the chain tolerance."""
    rng = np.random.default_rng(26)
    N, P, d = 30_000, 4, 4.0
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    bt = np.array([0.8, -0.6, 0.5, 2.2])
    y = rng.negative_binomial(d, d / (np.exp(X @ bt) + d)).astype(float)
    w, b, ds = gapi.nb_gibbs_df(y, X, np.zeros(P), 0.01 * np.eye(P), 12, 6, seed=8, d0=2.5, real_d=True)
    wo, bo, dso = loader.nb_gibbs_df(y, X, np.zeros(P), 0.01 * np.eye(P), 12, 6, seed=8, d0=2.5, real_d=True)
    assert len(set(ds)) > 2 and np.any(ds != np.floor(ds))
    close(ds, dso, 1e-12)
    close(b, bo, 1e-7); close(w, wo, 1e-7)


def test_em_matches_oracle_restatement(gapi):
    """logit.EM against the restatement of Logit::EM (Logit.hpp:488-554): same beta AND the same iteration count --
    the loop runs while max|beta - beta_old| > tol and iter < max_iter, and max_iter cuts it when tol is not met."""
    X, y, n, _ = synth_logit(20_000, 6, 9, binomial=True)
    for tol, max_iter in ((1e-6, 100), (1e-11, 200), (1e-14, 7)):
        out = gapi.logit_EM(y, X, n, tol=tol, max_iter=max_iter)
        bo, ito = loader.logit_em(y, X, n, tol=tol, max_iter=max_iter)
        assert out["iter"] == ito, (tol, max_iter, out["iter"], ito)
        assert np.allclose(out["beta"], bo, rtol=1e-9, atol=1e-12)
    assert gapi.logit_EM(y, X, n, tol=1e-14, max_iter=7)["iter"] == 7


def test_thinned_omega_chain(gapi):
    """bl_logit_gibbs_thin: beta every iteration, omega every w_every-th sampling iteration -- the rows of the
    unthinned chain, bit for bit."""
    X, y, n, _ = synth_logit(3000, 6, 5)
    P0 = 0.2 * np.eye(6)
    w, b = gapi.logit_gibbs(y, X, n, np.zeros(6), P0, 11, 4, seed=3)
    for every in (3, 5, 11, 20):
        wt, bt = gapi.logit_gibbs(y, X, n, np.zeros(6), P0, 11, 4, seed=3, w_every=every)
        assert np.array_equal(bt, b)
        assert wt.shape == ((11 + every - 1) // every, 3000) and np.array_equal(wt, w[::every])


def test_em_matches_newton_mode(gapi):
    """logit.EM (Logit.hpp:488-554): flat prior -> the MLE."""
    X, y, n, _ = synth_logit(4000, 6, 9, binomial=True)
    out = gapi.logit_EM(y, X, n, tol=1e-10, max_iter=200)
    b = np.zeros(6)
    for _ in range(60):
        p = 1 / (1 + np.exp(-X @ b))
        b = b + np.linalg.solve(X.T @ (X * (n * p * (1 - p))[:, None]), X.T @ (n * (y - p)))
    assert out["iter"] < 200
    assert np.allclose(out["beta"], b, rtol=1e-6, atol=1e-8)


def test_dropin_logit_and_mlogit_wrappers(gapi, engine):
    """The R-facing wrappers: combine first, then .C("gibbs") / .C("mult_gibbs")."""
    X, y, n, bt = synth_logit(5000, 4, 12)
    X = np.vstack([X, X[:50]]); y = np.r_[y, y[:50]]; n = np.r_[n, n[:50]]      # 50 duplicate rows
    engine.set_seed(5)
    out = gapi.logit(y, X, n, samp=300, burn=100, P0=0.01 * np.eye(4))
    assert out["X"].shape == (5000, 4) and out["w"].shape == (300, 5000) and out["beta"].shape == (300, 4)
    assert np.all(out["beta"][:, :-1] >= 0)                      # the reference's constrained draw
    assert np.max(np.abs(out["beta"].mean(0) - bt)) < 0.2
    engine.set_seed(5)
    again = gapi.logit(y, X, n, samp=300, burn=100, P0=0.01 * np.eye(4))
    assert np.array_equal(again["beta"], out["beta"])            # set.seed analogue
    assert gapi.logit(np.array([2.0, 0.0]), np.eye(2)) == -1     # check.parameters
    rng = np.random.default_rng(3)
    N, P, J = 3000, 3, 3
    Xm = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    B = rng.normal(0, 0.8, (P, J - 1))
    eta = np.c_[Xm @ B, np.zeros(N)]
    pr = np.exp(eta); pr /= pr.sum(1, keepdims=True)
    cat = (pr.cumsum(1) < rng.random(N)[:, None]).sum(1)
    Y = np.eye(J)[cat][:, :J - 1]
    om = gapi.mlogit(Y, Xm, samp=200, burn=100, P0=np.stack([0.01 * np.eye(P)] * (J - 1), axis=2))
    assert om["beta"].shape == (200, P, J - 1) and om["w"].shape == (200, N, J - 1)
    assert np.max(np.abs(om["beta"].mean(0) - B)) < 0.3


@pytest.mark.parametrize("constrained", [False, True])
@pytest.mark.parametrize("N,P,binomial", [(4099, 64, False), (200_001, 64, False), (2000, 6, True), (1777, 34, False), (50_000, 32, True),
                                          (97, 2, False)])
def test_one_pass_sweep_matches_two_pass_path(gapi, constrained, N, P, binomial):
    """k_logit_sweep (flag BL_GIBBS_ONE_PASS: psi, omega and X' Omega X from one TMA-staged read of X) against
    the default two-pass path (k_logit_psi_draw + k_gram_partial).  The two sum psi and the Gram in different
    orders (fp64, N terms), so the chains agree to rounding, not bit for bit: 1e-9 over 9 iterations.
    Covers every column-box count (P = 2, 6, 32, 34, 64), ragged last tiles, n_i in {1..5}, a grid of
    148 CTAs (N = 200 001) and grids smaller than the reduction's 128 entry runs."""
    X, y, n, _ = synth_logit(N, P, 99 + P, binomial)
    m0 = np.linspace(-0.1, 0.1, P)
    P0 = 0.3 * np.eye(P) + 0.01
    f = 0 if constrained else gapi.PLAIN_BETA
    w1, b1 = gapi.logit_gibbs(y, X, n, m0, P0, 6, 3, seed=15, flags=f | gapi.ONE_PASS)
    w2, b2 = gapi.logit_gibbs(y, X, n, m0, P0, 6, 3, seed=15, flags=f | gapi.TWO_PASS)
    close(b1, b2, 1e-9); close(w1, w2, 1e-9)
    assert np.all(w1 > 0)


@pytest.mark.parametrize("N,P,binomial", [(4099, 64, False), (2000, 6, True), (1777, 34, False)])
def test_fused_psi_draw_equals_two_kernel_path(gapi, N, P, binomial):
    """k_logit_psi_draw (psi = X beta and omega = PG(n, psi) in one pass over X) forms psi with the MMA
    order of k_xbeta_mma and draws with the sampler of k_devroye_refill: the chain -- omega and beta
    -- must carry the same bits as the two-kernel path (flag BL_GIBBS_UNFUSED), single chain and
    batched chains, ragged last trip (N not a multiple of 32) and n_i in {1..5} included."""
    X, y, n, _ = synth_logit(N, P, 77 + P, binomial)
    m0 = np.zeros(P)
    P0 = 0.3 * np.eye(P)
    w1, b1 = gapi.logit_gibbs(y, X, n, m0, P0, 6, 3, seed=5, flags=gapi.PLAIN_BETA | gapi.TWO_PASS)
    w2, b2 = gapi.logit_gibbs(y, X, n, m0, P0, 6, 3, seed=5, flags=gapi.PLAIN_BETA | gapi.UNFUSED)
    assert np.array_equal(w1, w2) and np.array_equal(b1, b2)
    assert np.all(w1 > 0) and np.all(np.isfinite(b1))
    C_ = 3
    Xc = np.stack([synth_logit(N, P, 200 + c, binomial)[0] for c in range(C_)])
    yc = np.stack([synth_logit(N, P, 200 + c, binomial)[1] for c in range(C_)])
    nc = np.stack([synth_logit(N, P, 200 + c, binomial)[2] for c in range(C_)])
    c1 = gapi.logit_chains(yc, Xc, nc, m0, P0, 5, 2, seed=9, flags=1)
    c2 = gapi.logit_chains(yc, Xc, nc, m0, P0, 5, 2, seed=9, flags=gapi.PLAIN_BETA | gapi.UNFUSED)
    assert np.array_equal(c1, c2)


def _combine(gapi, mode, y, X, n):
    import os
    os.environ["BAYESLOGIT_MERGE"] = mode
    try:
        return gapi.logit_combine(y, X, n) if np.ndim(y) == 1 else gapi.mlogit_combine(y, X, n)
    finally:
        os.environ.pop("BAYESLOGIT_MERGE", None)


@pytest.mark.parametrize("case", ["groups", "all_identical", "no_duplicates", "signed_zero", "nan_rows", "mlogit", "wide"])
def test_device_merge_equals_host_merge(gapi, case):
    """combine / mult_combine on the device (merge.cu: row hashes, stable radix sort by first occurrence, the
    reference's running weighted mean replayed per group; Logit.hpp:192-270, MultLogit.hpp:137-208) against the
    host merge: same rows in the same (first-occurrence) order, y and n BIT-identical.  Rows the hash cannot decide
    (NaN covariates never equal themselves) take the exact host path."""
    rng = np.random.default_rng(7)
    N, P, D = 120_000, 6, 900
    if case == "wide":
        N, P, D = 300_000, 64, 40_000
    base = rng.standard_normal((D, P))
    idx = rng.integers(0, D, N)
    X = base[idx].copy()
    n = rng.integers(1, 5, N).astype(float)
    y = rng.binomial(n.astype(int), 0.35) / n
    if case == "all_identical":
        X = np.ones((50_000, 1)); n = n[:50_000]; y = y[:50_000]
    elif case == "no_duplicates":
        X = rng.standard_normal((N, P))
    elif case == "signed_zero":
        X[:, 0] = np.where(rng.random(N) < 0.5, 0.0, -0.0)          # -0 == +0: such rows still merge
    elif case == "nan_rows":
        X[::1000, 2] = np.nan
    elif case == "mlogit":
        U = 3
        cat = rng.integers(0, U + 1, N)
        y = np.eye(U + 1)[cat][:, :U]
    dev = _combine(gapi, "device", y, X, n)
    host = _combine(gapi, "host", y, X, n)
    assert dev["X"].shape == host["X"].shape
    assert np.array_equal(dev["X"], host["X"], equal_nan=True)
    assert np.array_equal(dev["n"], host["n"]) and np.array_equal(dev["y"], host["y"])
    if case == "all_identical":
        assert dev["X"].shape[0] == 1 and dev["n"][0] == n.sum()
    if case == "no_duplicates":
        assert dev["X"].shape[0] == N
