"""CPU tests of the Gibbs restatements in oracle/gibbs_oracle.c (no reference output exists
to pin them -- "parity unpinned" -- so they are validated statistically), of the host-side
duplicate-row merge behind `combine` / `mult_combine`, and of the multi-rank sharding logic
under gloo."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth_logit(N, P, seed):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]          # intercept last
    bt = np.r_[np.abs(rng.normal(0, 0.5, P - 1)), -0.5]
    y = (rng.random(N) < 1 / (1 + np.exp(-X @ bt))).astype(float)
    return X, y, bt


def logit_map(X, y, n, P0, iters=50):
    b = np.zeros(X.shape[1])
    for _ in range(iters):
        p = 1 / (1 + np.exp(-X @ b))
        g = X.T @ (n * (y - p)) - P0 @ b
        H = X.T @ (X * (n * p * (1 - p))[:, None]) + P0
        b = b + np.linalg.solve(H, g)
    return b, np.linalg.inv(H)


def test_logit_gibbs_oracle_recovers_posterior():
    X, y, bt = synth_logit(3000, 5, 0)
    n = np.ones(len(y))
    P0 = 0.01 * np.eye(5)
    mode, cov = logit_map(X, y, n, P0)
    w, b = loader.logit_gibbs(y, X, n, np.zeros(5), P0, 600, 150, seed=1, constrained=False)
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(b.mean(0) - mode) < 0.35 * sd)
    assert np.all(np.abs(b.std(0) / sd - 1) < 0.25)
    assert w.shape == (600, 3000) and np.all(w > 0)
    # the constrained draw the reference actually calls keeps beta_j >= 0 for j < P-1
    w, b = loader.logit_gibbs(y, X, n, np.zeros(5), P0, 300, 100, seed=2, constrained=True)
    assert np.all(b[:, :-1] >= 0)
    assert np.all(np.abs(b.mean(0) - mode) < 0.6 * sd)


def test_logit_gibbs_slot_semantics():
    """Burn-in overwrites slot 0; sampling restarts from it (Logit.hpp:408-444,473-478)."""
    X, y, _ = synth_logit(500, 3, 3)
    n = np.ones(len(y))
    args = (y, X, n, np.zeros(3), np.eye(3))
    w1, b1 = loader.logit_gibbs(*args, 5, 0, seed=9, constrained=False)
    assert np.all(b1[0] != 0)
    # binomial counts: n > 1 draws sums of PG(1)
    n2 = np.full(len(y), 3.0)
    w2, _ = loader.logit_gibbs(y, X, n2, np.zeros(3), np.eye(3), 3, 2, seed=9, constrained=False)
    assert w2.mean() > 2 * w1.mean()


def test_mlogit_and_nb_oracle_recover_truth():
    rng = np.random.default_rng(4)
    N, P, J = 4000, 4, 3
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    B = rng.normal(0, 0.7, (P, J - 1))
    eta = np.c_[X @ B, np.zeros(N)]
    pr = np.exp(eta); pr /= pr.sum(1, keepdims=True)
    cat = (pr.cumsum(1) < rng.random(N)[:, None]).sum(1)
    Y = np.eye(J)[cat][:, :J - 1]
    P0 = np.stack([0.01 * np.eye(P)] * (J - 1), axis=2)
    w, b = loader.mlogit_gibbs(Y, X, np.ones(N), np.zeros((P, J - 1)), P0, 300, 100, seed=5)
    assert np.max(np.abs(b.mean(0) - B.T)) < 0.25
    d = 5.0
    bt = np.array([0.2, -0.1, 0.3, 1.5])
    mu = np.exp(X @ bt)
    yc = rng.negative_binomial(d, d / (mu + d)).astype(float)
    _, bn = loader.nb_gibbs(yc, X, d, np.zeros(P), 0.01 * np.eye(P), 300, seed=6)
    assert np.max(np.abs(bn[100:].mean(0) - bt)) < 0.1


def test_nb_oracle_with_dispersion_update_recovers_d_and_beta():
    """NB.PG.gibbs with draw.df (NBPG-logmean.R:36-113, NB-Shape.R:9-53): started at d = 1, the
    random-walk Metropolis step has to walk to the true dispersion and stay around it."""
    rng = np.random.default_rng(14)
    N, P, d = 3000, 3, 6.0
    X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
    bt = np.array([0.3, -0.2, 1.2])
    mu = np.exp(X @ bt)
    y = rng.negative_binomial(d, d / (mu + d)).astype(float)
    _, b, ds = loader.nb_gibbs_df(y, X, np.zeros(P), 0.01 * np.eye(P), 400, 200, seed=3)
    assert np.all(ds == np.floor(ds)) and ds.min() >= 1
    assert abs(ds.mean() - d) < 1.5
    assert np.max(np.abs(b.mean(0) - bt)) < 0.1


@pytest.fixture(scope="module")
def englib():
    from bayeslogit_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_native()
    return _lib.lib()


def test_combine_merges_duplicate_rows(englib):
    """`combine` (LogitWrapper.cpp:279-310 -> Logit::compress, Logit.hpp:192-270): rows with
    identical covariates merge in first-occurrence order, y = n-weighted mean, n = sum.
    Pure host code: runs without a GPU."""
    from bayeslogit_b200 import gibbs_api
    X = np.array([[1.0, 2.0], [0.0, 1.0], [1.0, 2.0], [3.0, 3.0], [0.0, 1.0], [1.0, 2.0]])
    y = np.array([1.0, 0.0, 0.0, 1.0, 1.0, 0.5])
    n = np.array([1.0, 2.0, 3.0, 1.0, 2.0, 4.0])
    out = gibbs_api.logit_combine(y, X, n)
    assert np.array_equal(out["X"], [[1.0, 2.0], [0.0, 1.0], [3.0, 3.0]])
    assert np.allclose(out["n"], [8.0, 4.0, 1.0])
    # the reference merges pairwise in sequence: ((1*1 + 3*0)/4 * 4 + 4*0.5)/8
    assert np.allclose(out["y"], [(1 * 1.0 + 3 * 0.0 + 4 * 0.5) / 8, (2 * 0.0 + 2 * 1.0) / 4, 1.0])
    # no duplicates: unchanged
    out2 = gibbs_api.logit_combine(y[:2], X[:2], n[:2])
    assert np.array_equal(out2["X"], X[:2]) and np.array_equal(out2["y"], y[:2])
    # multinomial flavour
    Y = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.0, 0.0]])
    outm = gibbs_api.mlogit_combine(Y, X, np.ones(6))
    assert outm["X"].shape == (3, 2) and np.allclose(outm["n"], [3.0, 2.0, 1.0])
    assert np.allclose(outm["y"][0], [1 / 3, 0.0]) and np.allclose(outm["y"][1], [0.5, 0.5])
    # validation like check.parameters (LogitWrapper.R:130-157)
    assert gibbs_api.logit_combine(np.array([2.0]), np.array([[1.0]]), np.array([1.0])) == -1


def test_sharded_gram_allreduce_gloo():
    """N > 1 host logic on CPU: two gloo ranks shard the observations, all-reduce the
    P*P + P block and end with the single-rank posterior precision and right-hand side."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29581")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29581",
                          os.path.join(ROOT, "tests", "_gloo_shard_worker.py")],
                         capture_output=True, text=True, env=env, timeout=300)
    assert "SHARD_OK" in out.stdout, out.stderr[-2000:]
