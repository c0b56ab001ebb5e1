"""CPU tests of the host-side mirror of the reference's R wrappers (bayeslogit_b200/api.py,
gibbs_api.py): the pre-validation of Code/R/LogitWrapper.R (messages and NA / -1 returns) happens
before any device work, and the duplicate-row merge (`combine`, `mult_combine`:
LogitWrapper.cpp:279-310, 376-409) is host code of the shared library -- neither needs a GPU.
Everything that would compute must fail loudly here (no CPU fallback)."""
import math

import numpy as np
import pytest

from bayeslogit_b200 import _lib, api, gibbs_api


def test_rpg_wrappers_validate_like_r(capsys):
    # rpg.gamma / rpg.devroye / rpg.alt / rpg.sp / rpg: LogitWrapper.R:15-22, 39-42, 59-62, 80-83, 107-110
    assert math.isnan(api.rpg_gamma(3, h=-1.0))
    assert math.isnan(api.rpg_gamma(3, h=1.0, trunc=0))
    assert math.isnan(api.rpg_devroye(3, n=-1))
    assert math.isnan(api.rpg_alt(3, h=0.5))
    assert math.isnan(api.rpg_sp(3, h=0.99))
    assert math.isnan(api.rpg(3, h=0.0))
    out = capsys.readouterr().out.splitlines()
    assert out == ["h must be greater than zero.", "trunc must be > 0.", "n must be greater than zero.",
                   "h must be >= 1.", "h must be >= 1.", "h must be > 0."]


def test_recycling_follows_r_array():
    # array(h, num) recycles a shorter vector (LogitWrapper.R:25-26)
    assert np.array_equal(api._recycle([1.0, 2.0], 5, np.float64), [1.0, 2.0, 1.0, 2.0, 1.0])
    assert np.array_equal(api._recycle(3, 4, np.int32), [3, 3, 3, 3])
    assert api._recycle([1, 2, 3], 2, np.float64).tolist() == [1.0, 2.0]


def test_check_parameters_messages(capsys):
    # check.parameters, LogitWrapper.R:130-157
    y, n = np.array([0.0, 1.0, 0.5]), np.ones(3)
    assert gibbs_api.check_parameters(y, n, np.zeros(2), np.eye(2), 3, 2, 10, 0)
    assert capsys.readouterr().out == ""
    assert not gibbs_api.check_parameters(np.array([0.0, 1.5, -0.1]), np.array([1.0, 0.0, 1.0]), np.zeros(3), np.eye(2), 3, 2, 0, -1)
    msg = capsys.readouterr().out
    for want in ("y must be >= 0.", "y is a proportion; it must be <= 1.", "n must be > 0.",
                 "col(X) != length(m0) 2 3", "samp must be > 0.", "burn must be >=0."):
        assert want in msg
    assert not gibbs_api.check_parameters(y, np.ones(2), np.zeros(2), np.eye(3), 3, 2, 1, 0)
    msg = capsys.readouterr().out
    assert "col(X) != row(P0) 2 3" in msg and "col(X) != col(P0) 2 3" in msg and "Dimensions do not conform" in msg


def test_mult_check_parameters_messages(capsys):
    # mult.check.parameters, LogitWrapper.R:294-321
    y = np.array([[0.2, 0.3], [0.6, 0.6]])
    X = np.ones((2, 3))
    assert not gibbs_api.mult_check_parameters(y, X, np.ones(2), np.zeros((3, 2)), np.zeros((3, 3, 2)), 1, 0)
    assert "y[i,] are proportions and must sum <= 1." in capsys.readouterr().out
    y[1] = [0.5, 0.5]
    assert gibbs_api.mult_check_parameters(y, X, np.ones(2), np.zeros((3, 2)), np.zeros((3, 3, 2)), 1, 0)
    assert not gibbs_api.mult_check_parameters(y, X, np.ones(2), np.zeros((2, 2)), np.zeros((3, 3, 1)), 1, 0)
    msg = capsys.readouterr().out
    assert "m.0 does not conform." in msg and "P.0 does not conform." in msg


def test_invalid_models_return_before_touching_the_device(capsys):
    X = np.ones((3, 2))
    _lib.lib().bl_clear_error()
    assert gibbs_api.logit(np.array([0.0, 2.0, 1.0]), X) == -1          # logit(): LogitWrapper.R:206-214
    assert gibbs_api.logit_EM(np.array([0.0, -1.0, 1.0]), X) == -1
    assert math.isnan(gibbs_api.mlogit(np.array([[0.9, 0.9]] * 3), X))  # mlogit(): LogitWrapper.R:369-377
    capsys.readouterr()
    assert _lib.lib().bl_last_error() in (None, b"")                    # nothing reached the library


def test_logit_combine_merges_duplicates_on_the_host():
    # Logit::compress (Logit.hpp:192-270): rows with equal covariates are merged, n summed,
    # y the n-weighted mean; first-occurrence order.  Host code: runs without a GPU.
    X = np.array([[1.0, 2.0], [0.0, 1.0], [1.0, 2.0], [3.0, 3.0], [0.0, 1.0], [1.0, 2.0]])
    y = np.array([1.0, 0.0, 0.0, 1.0, 1.0, 0.5])
    n = np.array([1.0, 2.0, 1.0, 1.0, 2.0, 2.0])
    out = gibbs_api.logit_combine(y, X, n)
    assert out["X"].tolist() == [[1.0, 2.0], [0.0, 1.0], [3.0, 3.0]]
    assert out["n"].tolist() == [4.0, 4.0, 1.0]
    assert np.allclose(out["y"], [(1 * 1 + 0 * 1 + 0.5 * 2) / 4, (0 * 2 + 1 * 2) / 4, 1.0])
    # successes are conserved
    assert math.isclose(float((out["y"] * out["n"]).sum()), float((y * n).sum()))


def test_mlogit_combine_merges_duplicates_on_the_host():
    # MultLogit::set_data merge (MultLogit.hpp:137-208)
    X = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0]])
    y = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])
    out = gibbs_api.mlogit_combine(y, X, np.array([1.0, 1.0, 3.0]))
    assert out["X"].tolist() == [[1.0, 0.0], [0.0, 1.0]]
    assert out["n"].tolist() == [4.0, 1.0]
    assert np.allclose(out["y"], [[0.25, 0.0], [0.0, 1.0]])


def test_compute_entry_points_fail_loudly_without_a_gpu(capsys):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        api.rpg_devroye(4, 1, 0.3)
    with pytest.raises(RuntimeError):
        gibbs_api.logit(np.array([0.0, 1.0, 1.0]), np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]), samp=2, burn=0)


def _schedule(num):
    import ctypes as C
    buf = (C.c_int64 * 4096)()
    k = _lib.lib().bl_probe_pipeline_schedule(int(num), C.cast(buf, C.c_void_p), 4096)
    return list(buf[:k]) if k >= 0 else None


def test_host_pipeline_schedule():
    """Chunk schedule of the host-pointer entry points (capi.cu run_host): covers the batch exactly, full-size
    chunks of 4M observations, and for batches of >= 4 chunks a ramp K/8, K/4, K/2 at both ends."""
    K = 1 << 22
    assert _schedule(-1) is None
    assert _schedule(0) == []
    assert _schedule(1) == [1]
    assert _schedule(K) == [K]
    assert _schedule(K + 5) == [K, 5]
    assert _schedule(4 * K - 1) == [K, K, K, K - 1]                     # below the ramp threshold
    ramp = [K // 8, K // 4, K // 2]
    s = _schedule(4 * K)
    assert s[:3] == ramp and s[-3:] == ramp[::-1] and sum(s) == 4 * K
    for num in (100_000_000, 34 * (1 << 20) + 12345, 2**31 - 1, 5 * K + 1):
        s = _schedule(num)
        assert sum(s) == num and all(0 < c <= K for c in s)
        assert s[:3] == ramp and s[-3:] == ramp[::-1]
        body = s[3:-3]
        assert all(c == K for c in body[:-1])                           # only the last body chunk may be short


def test_logit_combine_at_scale_matches_a_groupby():
    """the reference's merge is O(N^2 P) over a std::list (Logit.hpp:192-270); here it is
    one hash pass.  200k rows drawn from 3000 distinct covariate vectors: first-occurrence order, trial counts
    and successes per group equal a numpy group-by."""
    rng = np.random.default_rng(0)
    N, P, D = 200_000, 16, 3000
    base = rng.standard_normal((D, P))
    idx = rng.integers(0, D, N)
    X = base[idx]
    n = rng.integers(1, 4, N).astype(float)
    y = rng.binomial(n.astype(int), 0.4) / n
    out = gibbs_api.logit_combine(y, X, n)
    first = np.sort(np.unique(idx, return_index=True)[1])
    assert np.array_equal(out["X"], X[first])
    trials = np.bincount(idx, weights=n, minlength=D)[idx[first]]
    succ = np.bincount(idx, weights=y * n, minlength=D)[idx[first]]
    assert np.allclose(out["n"], trials, rtol=0, atol=0)
    assert np.allclose(out["y"] * out["n"], succ, rtol=1e-12, atol=1e-9)
