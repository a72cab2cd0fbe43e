"""ctypes binding of oracle/_build/libhill_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhill_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.phf_oracle_log_ndtr.restype = C.c_double
        L.phf_oracle_log_ndtr.argtypes = [C.c_double]
        L.phf_oracle_ndtr.restype = C.c_double
        L.phf_oracle_ndtr.argtypes = [C.c_double]
        L.phf_oracle_log_target_batch.restype = None
        L.phf_oracle_log_target_batch.argtypes = [C.c_int, C.c_int, _dp, _dp, _u8p, C.c_int, _dp, _dp, C.c_double,
                                                  _dp, _dp]
        L.phf_oracle_hier_log_target_batch.restype = None
        L.phf_oracle_hier_log_target_batch.argtypes = [C.c_int, _ip, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.phf_oracle_philox.restype = None
        L.phf_oracle_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                        np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
        L.phf_oracle_draw.restype = None
        L.phf_oracle_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _dp, _dp]
        L.phf_oracle_am_single.restype = C.c_int
        L.phf_oracle_am_single.argtypes = [C.c_int, C.c_int, _dp, _dp, _u8p, C.c_double, C.c_double, _dp,
                                           C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64,
                                           C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.phf_oracle_am_hier.restype = C.c_int
        L.phf_oracle_am_hier.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.phf_oracle_am_single_many.restype = C.c_int
        L.phf_oracle_am_single_many.argtypes = [
            C.c_int, _ip, _ip, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), _dp, _dp, _u8p, _dp,
            _dp, _dp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), C.c_uint32, C.c_uint32,
            C.c_uint32, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"), C.c_int, C.c_uint64,
            np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"), C.c_int]
        _lib = L
    return _lib


def classify(responses):
    """0: y == 0, 1: y == 100, 2: 0 < y < 100, 3: in no mask (PyHillFit.py:675-677)."""
    r = np.asarray(responses, dtype=np.float64)
    cls = np.full(r.shape, 3, dtype=np.uint8)
    cls[r == 0] = 0
    cls[r == 100] = 1
    cls[(0 < r) & (r < 100)] = 2
    return cls


def log_target_batch(model, concs, responses, theta, t, pi_bit, cls=None):
    concs = np.ascontiguousarray(concs, dtype=np.float64)
    responses = np.ascontiguousarray(responses, dtype=np.float64)
    cls = classify(responses) if cls is None else np.ascontiguousarray(cls, dtype=np.uint8)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=np.float64), (theta.shape[0],)))
    out = np.empty(theta.shape[0])
    ll1 = np.empty(theta.shape[0])
    lib().phf_oracle_log_target_batch(model, len(concs), concs, responses, cls, theta.shape[0], theta, t,
                                      float(pi_bit), out, ll1)
    return out, ll1


def hier_pack(experiments):
    off = np.zeros(len(experiments) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(e) for e in experiments])
    conc = np.ascontiguousarray(np.concatenate([np.asarray(e)[:, 0] for e in experiments]), dtype=np.float64)
    y = np.ascontiguousarray(np.concatenate([np.asarray(e)[:, 1] for e in experiments]), dtype=np.float64)
    return off, conc, y


def hier_log_target_batch(experiments, theta, shapes, scales, locs):
    off, conc, y = hier_pack(experiments)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.empty(theta.shape[0])
    lib().phf_oracle_hier_log_target_batch(len(experiments), off, conc, y, theta.shape[0], theta,
                                           np.ascontiguousarray(shapes, dtype=np.float64),
                                           np.ascontiguousarray(scales, dtype=np.float64),
                                           np.ascontiguousarray(locs, dtype=np.float64), out)
    return out


def state_size(d):
    return 2 * d + d * d + 5


def make_state(theta0, lt, ll1, cov0):
    d = len(theta0)
    s = np.zeros(state_size(d))
    s[:d] = theta0
    s[d] = lt
    s[d + 1] = ll1
    s[d + 2:2 * d + 2] = theta0
    cov = np.array(cov0, dtype=np.float64).reshape(d, d)
    dg = np.diag(cov).copy()
    cov[np.diag_indices(d)] = np.where(dg > 0, dg, 1e-60)   # PHF_COV0_DIAG_FLOOR (hill_oracle.COV0_DIAG_FLOOR)
    s[2 * d + 2:2 * d + 2 + d * d] = cov.reshape(-1)
    return s


def am_single(model, concs, responses, temperature, pi_bit, state, t0, iters, thinning, adapt_when, reset_mean,
              seed, chain_id, burn=0xFFFFFFFF, want_chain=True, cls=None):
    concs = np.ascontiguousarray(concs, dtype=np.float64)
    responses = np.ascontiguousarray(responses, dtype=np.float64)
    cls = classify(responses) if cls is None else cls
    d = 2 if model == 1 else 3
    row0 = t0 // thinning + 1
    nrows = (t0 + iters) // thinning - t0 // thinning
    chain = np.zeros((nrows, d + 1)) if want_chain else None
    rc = lib().phf_oracle_am_single(model, len(concs), concs, responses, cls, float(temperature), float(pi_bit),
                                    state, t0, iters, thinning, adapt_when, int(reset_mean), seed, chain_id, row0,
                                    burn, chain.ctypes.data if want_chain else None)
    if rc:
        raise RuntimeError("oracle AM failed (non-PD covariance)")
    return chain


def am_hier(experiments, shapes, scales, locs, state, t0, iters, thinning, adapt_when, seed, chain_id):
    off, conc, y = hier_pack(experiments)
    d = 5 + 2 * len(experiments)
    row0 = t0 // thinning + 1
    nrows = (t0 + iters) // thinning - t0 // thinning
    chain = np.zeros((nrows, d + 1))
    rc = lib().phf_oracle_am_hier(len(experiments), off, conc, y, np.ascontiguousarray(shapes, dtype=np.float64),
                                  np.ascontiguousarray(scales, dtype=np.float64),
                                  np.ascontiguousarray(locs, dtype=np.float64), state, t0, iters, thinning,
                                  adapt_when, seed, chain_id, row0, chain.ctypes.data)
    if rc:
        raise RuntimeError("oracle AM failed (non-PD covariance)")
    return chain
