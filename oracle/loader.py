"""oracle/loader.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for the two CPU oracle libraries (see oracle/batch.h):

  kind="port"       oracle/_build/libpg_oracle.so   plain-C restatement
  kind="reference"  oracle/_ref/libpg_ref.so        reference sources compiled in place

May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by bayeslogit_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {
    "port": os.path.join(HERE, "_build", "libpg_oracle.so"),
    "reference": os.path.join(HERE, "_ref", "libpg_ref.so"),
}
TRACE_W = 6
MODE_TAPE, MODE_PHILOX = 0, 1


class Stream(C.Structure):
    _fields_ = [("mode", C.c_int32), ("lu", C.c_int32), ("le", C.c_int32),
                ("ln", C.c_int32), ("lg", C.c_int32),
                ("tu", C.c_void_p), ("te", C.c_void_p), ("tn", C.c_void_p), ("tg", C.c_void_p),
                ("seed", C.c_uint64), ("obs0", C.c_uint64), ("call_id", C.c_uint32)]


def build(targets=("port", "ref")):
    """(Re)build the oracle libraries with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", HERE, *targets], check=True)


def available(kind):
    return os.path.exists(PATHS[kind])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, kind="port"):
        path = PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle`")
        self.lib = lib = C.CDLL(path)
        lib.pgb_kind.restype = C.c_char_p
        assert lib.pgb_kind().decode() == kind
        self.kind = kind
        dp, ip, vp = C.c_void_p, C.c_void_p, C.c_void_p
        sp = C.POINTER(Stream)
        lib.pgb_rpg_devroye.argtypes = [dp, ip, dp, C.c_int, sp, vp, C.c_int]
        lib.pgb_rpg_gamma.argtypes = [dp, dp, dp, C.c_int, C.c_int, sp, vp, C.c_int]
        lib.pgb_rpg_alt.argtypes = [dp, dp, dp, C.c_int, sp, vp, C.c_int]
        lib.pgb_rpg_sp.argtypes = [dp, dp, dp, C.c_int, ip, sp, vp, C.c_int]
        lib.pgb_rpg_hybrid.argtypes = [dp, dp, dp, C.c_int, sp, vp, C.c_int]
        for f in (lib.pgb_pg_m1, lib.pgb_pg_m2):
            f.argtypes = [C.c_double, C.c_double]
            f.restype = C.c_double
        lib.pgb_v_eval.argtypes = [C.c_double]
        lib.pgb_v_eval.restype = C.c_double
        if kind == "port":   # the Gibbs restatements exist only in the plain-C port
            vp, ci, u64 = C.c_void_p, C.c_int, C.c_uint64
            lib.pgb_logit_gibbs.argtypes = [vp] * 7 + [ci, ci, ci, ci, u64, ci, ci]
            lib.pgb_mlogit_gibbs.argtypes = [vp] * 7 + [ci, ci, ci, ci, ci, u64, ci]
            lib.pgb_nb_gibbs.argtypes = [vp, vp, vp, vp, C.c_double, vp, vp, ci, ci, ci, u64, ci]
            lib.pgb_nb_gibbs_df.argtypes = [vp, vp, vp, vp, vp, C.c_double, vp, vp, ci, ci, ci, ci, u64, ci]
            lib.pgb_nb_gibbs_dfreal.argtypes = [vp, vp, vp, vp, vp, C.c_double, vp, vp, ci, ci, ci, ci, u64, ci]
            lib.pgb_logit_em.argtypes = [vp, vp, vp, vp, ci, ci, C.c_double, ci, ci]

    # -- stream helpers ------------------------------------------------------
    @staticmethod
    def _stream(num, seed=None, obs0=0, call_id=0, tape=None):
        st = Stream()
        keep = []
        if tape is not None:
            st.mode = MODE_TAPE
            for k in "ueng":
                a = tape.get(k)
                if a is None:
                    setattr(st, "l" + k, 0)
                    setattr(st, "t" + k, None)
                else:
                    a = _f64(a)
                    assert a.ndim == 2 and a.shape[0] == num, (k, a.shape, num)
                    keep.append(a)
                    setattr(st, "l" + k, a.shape[1])
                    setattr(st, "t" + k, a.ctypes.data)
        else:
            st.mode = MODE_PHILOX
            st.seed = int(seed)
            st.obs0 = int(obs0)
            st.call_id = int(call_id)
        return st, keep

    def _run(self, fn, args_before, num, extra, seed, obs0, call_id, tape, trace, nthreads):
        st, keep = self._stream(num, seed, obs0, call_id, tape)
        x = np.zeros(num, dtype=np.float64)
        tr = np.zeros((num, TRACE_W), dtype=np.int32) if trace else None
        fn(x.ctypes.data, *[a.ctypes.data for a in args_before], num, *extra, C.byref(st),
           tr.ctypes.data if trace else None, nthreads)
        del keep
        return (x, tr) if trace else x

    # -- batch entry points (reference: LogitWrapper.cpp:39-167) ---------------
    def rpg_devroye(self, n, z, seed=0, obs0=0, call_id=0, tape=None, trace=False, nthreads=1):
        n = np.ascontiguousarray(n, dtype=np.int32)
        z = _f64(z)
        return self._run(self.lib.pgb_rpg_devroye, [n, z], len(z), [], seed, obs0, call_id,
                         tape, trace, nthreads)

    def rpg_gamma(self, n, z, trunc=200, seed=0, obs0=0, call_id=0, tape=None, trace=False,
                  nthreads=1):
        n, z = _f64(n), _f64(z)
        return self._run(self.lib.pgb_rpg_gamma, [n, z], len(z), [int(trunc)], seed, obs0,
                         call_id, tape, trace, nthreads)

    def rpg_alt(self, h, z, seed=0, obs0=0, call_id=0, tape=None, trace=False, nthreads=1):
        h, z = _f64(h), _f64(z)
        return self._run(self.lib.pgb_rpg_alt, [h, z], len(z), [], seed, obs0, call_id,
                         tape, trace, nthreads)

    def rpg_sp(self, h, z, seed=0, obs0=0, call_id=0, tape=None, trace=False, nthreads=1):
        h, z = _f64(h), _f64(z)
        it = np.zeros(len(z), dtype=np.int32)
        out = self._run(self.lib.pgb_rpg_sp, [h, z], len(z), [it.ctypes.data], seed, obs0,
                        call_id, tape, trace, nthreads)
        return (*out, it) if trace else (out, it)

    def rpg_hybrid(self, h, z, seed=0, obs0=0, call_id=0, tape=None, trace=False, nthreads=1):
        h, z = _f64(h), _f64(z)
        return self._run(self.lib.pgb_rpg_hybrid, [h, z], len(z), [], seed, obs0, call_id,
                         tape, trace, nthreads)

    def pg_m1(self, b, z):
        return self.lib.pgb_pg_m1(float(b), float(z))

    def pg_m2(self, b, z):
        return self.lib.pgb_pg_m2(float(b), float(z))

    def v_eval(self, y):
        return self.lib.pgb_v_eval(float(y))


def _gibbs_port():
    if not available("port"):
        build(("port",))
    return Oracle("port")


def logit_gibbs(y, X, n, m0, P0, samp, burn, seed, constrained=True, nthreads=0):
    """CPU restatement of Logit::gibbs (oracle/gibbs_oracle.c).  X: N x P row-major.
    Returns (w [samp x N], beta [samp x P])."""
    O = _gibbs_port()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    y, n, m0 = _f64(y).ravel(), _f64(n).ravel(), _f64(m0).ravel()
    P0c = np.asfortranarray(_f64(P0))
    w, beta = np.zeros((samp, N)), np.zeros((samp, P))
    st = O.lib.pgb_logit_gibbs(w.ctypes.data, beta.ctypes.data, y.ctypes.data, X.ctypes.data,
                               n.ctypes.data, m0.ctypes.data, P0c.ctypes.data, N, P, samp, burn,
                               int(seed), 1 if constrained else 0, nthreads)
    if st:
        raise RuntimeError("oracle logit_gibbs: precision not positive definite")
    return w, beta


def mlogit_gibbs(y, X, n, m0, P0, samp, burn, seed, nthreads=0):
    """CPU restatement of MultLogit::gibbs.  y: N x (J-1).  Returns (w [samp x (J-1) x N],
    beta [samp x (J-1) x P])."""
    O = _gibbs_port()
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, P = X.shape
    U = y.shape[1]
    n = _f64(n).ravel()
    m0c, P0c = np.asfortranarray(_f64(m0)), np.asfortranarray(_f64(P0))
    w, beta = np.zeros((samp, U, N)), np.zeros((samp, U, P))
    st = O.lib.pgb_mlogit_gibbs(w.ctypes.data, beta.ctypes.data, y.ctypes.data, X.ctypes.data,
                                n.ctypes.data, m0c.ctypes.data, P0c.ctypes.data, N, P, U + 1, samp, burn,
                                int(seed), nthreads)
    if st:
        raise RuntimeError("oracle mlogit_gibbs: precision not positive definite")
    return w, beta


def nb_gibbs(y, X, d, m0, P0, samp, seed, nthreads=0):
    """CPU restatement of the NB sweep with fixed d.  Returns (w_last [N], beta [samp x P])."""
    O = _gibbs_port()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    y, m0 = _f64(y).ravel(), _f64(m0).ravel()
    P0c = np.asfortranarray(_f64(P0))
    w, beta = np.zeros(N), np.zeros((samp, P))
    st = O.lib.pgb_nb_gibbs(w.ctypes.data, beta.ctypes.data, y.ctypes.data, X.ctypes.data, float(d),
                            m0.ctypes.data, P0c.ctypes.data, N, P, samp, int(seed), nthreads)
    if st:
        raise RuntimeError("oracle nb_gibbs: precision not positive definite")
    return w, beta


def nb_gibbs_df(y, X, m0, P0, samp, burn, seed, d0=1.0, nthreads=0, real_d=False):
    """CPU restatement of NB.PG.gibbs with draw.df (real_d: with draw.df.real.mean, NB-Shape.R:86-96).
    Returns (w_last [N], beta [samp x P], d [samp])."""
    O = _gibbs_port()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    y, m0 = _f64(y).ravel(), _f64(m0).ravel()
    P0c = np.asfortranarray(_f64(P0))
    w, beta, d = np.zeros(N), np.zeros((samp, P)), np.zeros(samp)
    fn = O.lib.pgb_nb_gibbs_dfreal if real_d else O.lib.pgb_nb_gibbs_df
    st = fn(w.ctypes.data, beta.ctypes.data, d.ctypes.data, y.ctypes.data, X.ctypes.data,
            float(d0), m0.ctypes.data, P0c.ctypes.data, N, P, samp, burn, int(seed), nthreads)
    if st:
        raise RuntimeError("oracle nb_gibbs_df: precision not positive definite")
    return w, beta, d


def logit_em(y, X, n, tol=1e-9, max_iter=100, nthreads=0):
    """CPU restatement of Logit::EM (Logit.hpp:488-554).  Returns (beta [P], iterations)."""
    O = _gibbs_port()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, P = X.shape
    y, n = _f64(y).ravel(), _f64(n).ravel()
    beta = np.zeros(P)
    it = O.lib.pgb_logit_em(beta.ctypes.data, y.ctypes.data, X.ctypes.data, n.ctypes.data, N, P, float(tol),
                            int(max_iter), nthreads)
    if it < 0:
        raise RuntimeError("oracle logit_em: X' Omega X not positive definite")
    return beta, it


def make_tape(num, lu=0, le=0, ln=0, lg=0, g_shape=None, seed=0):
    """Random variate tape: U(0,1), Exp(1), N(0,1) and Gamma(g_shape[i],1) rows."""
    rng = np.random.default_rng(seed)
    tape = {}
    if lu:
        tape["u"] = rng.random((num, lu))
    if le:
        tape["e"] = rng.standard_exponential((num, le))
    if ln:
        tape["n"] = rng.standard_normal((num, ln))
    if lg:
        shape = np.broadcast_to(np.asarray(g_shape, dtype=np.float64), (num,))
        tape["g"] = rng.standard_gamma(shape[:, None], size=(num, lg))
    return tape
