"""Sanity across P (small, odd, 32 < P < 64): logit chains, both beta draws, against the oracle; the constrained draw
with and without the speculative sweeps.  Run on a GPU box: python tools/check_tiny_p.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslogit_b200 import gibbs_api as gapi   # noqa: E402
from oracle import loader                        # noqa: E402

bad = 0
for P in (1, 2, 3, 9, 33, 40, 48, 63, 64):
    for constrained in (True, False):
        rng = np.random.default_rng(P)
        N = 100 * P + 1000
        X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)] if P > 1 else np.ones((N, 1))
        # coefficients the constraint (beta_j >= 0, j < P - 1) does not bind for: with binding constraints the truncation
        # windows are ulps wide and the chain map amplifies last-bit differences of erfc (tests/test_gpu_gibbs.py)
        bt = np.r_[np.abs(rng.normal(0, 0.4, P - 1)), -0.5] if P > 1 else np.array([-0.5])
        y = (rng.random(N) < 1 / (1 + np.exp(-X @ bt))).astype(float)
        m0, P0 = np.zeros(P), 0.1 * np.eye(P)
        fl = 0 if constrained else gapi.PLAIN_BETA
        wo, bo = loader.logit_gibbs(y, X, np.ones(N), m0, P0, 6, 3, seed=5, constrained=constrained)
        for nospec in ((False, True) if constrained else (False,)):
            if nospec:
                os.environ["BL_BETA_NO_SPEC"] = "1"
            w, b = gapi.logit_gibbs(y, X, np.ones(N), m0, P0, 6, 3, seed=5, flags=fl)
            os.environ.pop("BL_BETA_NO_SPEC", None)
            err = np.max(np.abs(b - bo) / np.maximum(np.abs(bo), 1.0))
            first = np.max(np.abs(b[0] - bo[0]) / np.maximum(np.abs(bo[0]), 1.0))
            print(f"P={P} constrained={constrained} nospec={nospec}: max rel err beta {err:.3e} (first kept iteration {first:.3e})")
            bad += err > 1e-8
print("TINY_P_OK" if not bad else f"TINY_P_FAILED ({bad})")
