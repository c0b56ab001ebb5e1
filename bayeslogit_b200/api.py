"""Host-side mirror of the reference's R front end for the Polya-Gamma path.

Names, argument meaning and error behaviour follow Code/R/LogitWrapper.R:
    rpg_gamma   <- rpg.gamma    (LogitWrapper.R:12-32)
    rpg_devroye <- rpg.devroye  (:36-52)
    rpg_alt     <- rpg.alt      (:56-72)
    rpg_sp      <- rpg.sp       (:77-100)
    rpg         <- rpg          (:104-121)
Each wrapper validates like its R counterpart (printing the same message and
returning NaN where R returns NA), recycles h/z to length `num`, and makes the
same C call R makes through .C(): host numpy buffers in, draws out.
"""
import ctypes as C

import numpy as np

from . import _lib

NA = float("nan")


def set_seed(seed):
    """Engine analogue of R's set.seed(): fixes the Philox key and resets the call counter."""
    _lib.lib().bl_set_seed(int(seed) & 0xFFFFFFFFFFFFFFFF)


def set_device(device):
    _lib.check(_lib.lib().bl_set_device(int(device)))


def _recycle(a, num, dtype):
    a = np.asarray(a, dtype=dtype).ravel()
    if a.size != num:
        a = np.resize(a, num)  # R's array(a, num) recycling
    return np.ascontiguousarray(a)


def _ptr(a):
    return a.ctypes.data


def _cint(v):
    return C.byref(C.c_int(int(v)))


def rpg_gamma(num=1, h=1, z=0.0, trunc=200):
    if np.any(np.asarray(h) < 0):
        print("h must be greater than zero.")
        return NA
    if trunc < 1:
        print("trunc must be > 0.")
        return NA
    x = np.zeros(num)
    h = _recycle(h, num, np.float64)
    z = _recycle(z, num, np.float64)
    _lib.lib().rpg_gamma(_ptr(x), _ptr(h), _ptr(z), _cint(num), _cint(trunc))
    _lib.check()
    return x


def rpg_devroye(num=1, n=1, z=0.0):
    if np.any(np.asarray(n) < 0):
        print("n must be greater than zero.")
        return NA
    x = np.zeros(num)
    n = _recycle(np.asarray(n).astype(np.int64), num, np.int32)
    z = _recycle(z, num, np.float64)
    _lib.lib().rpg_devroye(_ptr(x), _ptr(n), _ptr(z), _cint(num))
    _lib.check()
    return x


def rpg_alt(num=1, h=1, z=0.0):
    if np.any(np.asarray(h) < 1):
        print("h must be >= 1.")
        return NA
    x = np.zeros(num)
    h = _recycle(h, num, np.float64)
    z = _recycle(z, num, np.float64)
    _lib.lib().rpg_alt(_ptr(x), _ptr(h), _ptr(z), _cint(num))
    _lib.check()
    return x


def rpg_sp(num=1, h=1, z=0.0, track_iter=False):
    if np.any(np.asarray(h) < 1):
        print("h must be >= 1.")
        return NA
    x = np.zeros(num)
    it = np.zeros(num, dtype=np.int32)
    h = _recycle(h, num, np.float64)
    z = _recycle(z, num, np.float64)
    _lib.lib().rpg_sp(_ptr(x), _ptr(h), _ptr(z), _cint(num), _ptr(it))
    _lib.check()
    return {"samp": x, "iter": it} if track_iter else x


def rpg(num=1, h=1, z=0.0):
    if np.any(np.asarray(h) <= 0):
        print("h must be > 0.")
        return NA
    x = np.zeros(num)
    h = _recycle(h, num, np.float64)
    z = _recycle(z, num, np.float64)
    _lib.lib().rpg_hybrid(_ptr(x), _ptr(h), _ptr(z), _cint(num))
    _lib.check()
    return x


# ---------------------------------------------------------------------------
# Engine extensions (no R counterpart): explicit stream identity, tapes, probes
# ---------------------------------------------------------------------------

_SEEDED = {"devroye": ("bl_rpg_devroye_seeded", np.int32), "gamma": ("bl_rpg_gamma_seeded", np.float64),
           "alt": ("bl_rpg_alt_seeded", np.float64), "sp": ("bl_rpg_sp_seeded", np.float64),
           "hybrid": ("bl_rpg_hybrid_seeded", np.float64)}
_TAPE = {k: v[0].replace("_seeded", "_tape") for k, v in _SEEDED.items()}
_TAPE["devroye_plain"] = "bl_rpg_devroye_plain_tape"   # all-fp64 path without the fp32 decision filters


def rpg_seeded(method, shape, z, seed, call_id=0, obs0=0, trunc=200):
    """Draw with an explicit (seed, call_id, obs0) stream identity; host arrays."""
    fn, dt = _SEEDED[method]
    z = np.ascontiguousarray(z, dtype=np.float64)
    shape = np.ascontiguousarray(shape, dtype=dt)
    num = z.size
    x = np.zeros(num)
    args = [_ptr(x), _ptr(shape), _ptr(z), num]
    it = None
    if method == "gamma":
        args.append(int(trunc))
    if method == "sp":
        it = np.zeros(num, dtype=np.int32)
        args.append(_ptr(it))
    st = getattr(_lib.lib(), fn)(*args, int(seed), int(call_id), int(obs0))
    _lib.check(st)
    return (x, it) if method == "sp" else x


def rpg_tape(method, shape, z, tape, trunc=200, trace=True):
    """Draw from injected variate tapes (dict with optional 'u','e','n','g' [num x L] arrays)."""
    fn = _TAPE[method]
    dt = np.int32 if method.startswith("devroye") else np.float64
    z = np.ascontiguousarray(z, dtype=np.float64)
    shape = np.ascontiguousarray(shape, dtype=dt)
    num = z.size
    x = np.zeros(num)
    t = _lib.Tape()
    keep = []
    for k in "ueng":
        a = tape.get(k)
        if a is None:
            setattr(t, "t" + k, None)
            setattr(t, "l" + k, 0)
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.ndim == 2 and a.shape[0] == num
            keep.append(a)
            setattr(t, "t" + k, a.ctypes.data)
            setattr(t, "l" + k, a.shape[1])
    tr = np.zeros((num, _lib.TRACE_W), dtype=np.int32) if trace else None
    args = [_ptr(x), _ptr(shape), _ptr(z), num]
    it = None
    if method == "gamma":
        args.append(int(trunc))
    if method == "sp":
        it = np.zeros(num, dtype=np.int32)
        args.append(_ptr(it))
    st = getattr(_lib.lib(), fn)(*args, C.byref(t), _ptr(tr) if trace else None)
    _lib.check(st)
    out = [x]
    if trace:
        out.append(tr)
    if it is not None:
        out.append(it)
    return tuple(out) if len(out) > 1 else x


def pg_moments(b, z):
    b = np.ascontiguousarray(b, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    m1, m2 = np.zeros(b.size), np.zeros(b.size)
    _lib.check(_lib.lib().bl_probe_pg_moments(_ptr(m1), _ptr(m2), _ptr(b), _ptr(z), b.size))
    return m1, m2


def v_eval(y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    v = np.zeros(y.size)
    _lib.check(_lib.lib().bl_probe_v_eval(_ptr(v), _ptr(y), y.size))
    return v


SPECFUN = {"p_norm": 0, "log_p_norm": 1, "p_gamma_rate": 2, "p_igauss": 3, "lgamma": 4, "tgamma": 5,
           "dev_right_mass": 6, "dev_coef": 7, "upper_gamma_cf": 8, "p_igauss_direct": 9, "sp_log_cos_rt": 10,
           "sp_v_table": 11, "sp_pl": 12, "sp_pl_estimate": 13,
           "alt_pr": 14, "alt_pr_estimate": 15}


def specfun(which, a, b=None, c=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = a if b is None else np.ascontiguousarray(np.broadcast_to(b, a.shape), dtype=np.float64)
    c = a if c is None else np.ascontiguousarray(np.broadcast_to(c, a.shape), dtype=np.float64)
    out = np.zeros(a.size)
    _lib.check(_lib.lib().bl_probe_specfun(_ptr(out), SPECFUN[which], _ptr(a), _ptr(b), _ptr(c), a.size))
    return out


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _lib.check(_lib.lib().bl_probe_philox(_ptr(out), _ptr(ctr), _ptr(key)))
    return out
