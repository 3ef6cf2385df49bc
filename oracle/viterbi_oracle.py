"""Oracle (test infrastructure): Viterbi smoothing as the reference calls it.

Restates reference ``inaSpeechSegmenter/pyannote_viterbi.py:118-224`` (plain
max-sum DP, uniform prior, first-max tie-break) and the transition / emission
helpers of ``inaSpeechSegmenter/viterbi_utils.py:29-49``.

Two implementations of the same order of floating-point operations:
``viterbi_numpy`` (pure numpy/Python loop, small cases) and ``viterbi_c``
(``viterbi_oracle.c`` through ctypes, long sequences).  Both are pinned
against the real ``viterbi_decoding`` by ``tests/golden/make_golden.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(['make', '-s', '-C', _HERE], check=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        for name, etype in (('oracle_viterbi', ctypes.c_double), ('oracle_viterbi_f32', ctypes.c_float)):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.POINTER(etype), ctypes.POINTER(ctypes.c_double),
                           ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
        _LIB = lib
    return _LIB


# ---- viterbi_utils.py restated -------------------------------------------

def pred2logemission(pred, eps=1e-10):
    """viterbi_utils.py:29-34 -- 2-state log emissions from a boolean track."""
    pred = np.asarray(pred)
    em = np.full((len(pred), 2), eps)
    em[pred == 0, 0] = 1 - eps
    em[pred == 1, 1] = 1 - eps
    return np.log(em)


def log_trans_exp(exp, cost0=0, cost1=0):
    """viterbi_utils.py:36-42."""
    c = -exp * np.log(10)
    return np.array([[cost0, c], [c, cost1]], dtype=np.float64)


def diag_trans_exp(exp, dim):
    """viterbi_utils.py:44-49."""
    a = np.full((dim, dim), -exp * np.log(10))
    a[np.arange(dim), np.arange(dim)] = 0
    return a


# ---- the DP -----------------------------------------------------------------

def viterbi_numpy(emission, transition):
    """Straight numpy restatement; returns float array like the reference
    (``_update_states`` builds ``np.empty(states.shape)``, :106)."""
    emission = np.asarray(emission)
    T, K = emission.shape
    V = emission[0] + np.log(np.ones(K) / K)
    P = np.zeros((T, K), dtype=np.int64)
    cols = np.arange(K)
    for t in range(1, T):
        cand = V[:, None] + transition           # cand[k, j] = V[k] + A[k, j]
        P[t] = np.argmax(cand, axis=0)
        V = emission[t] + cand[P[t], cols]
    X = np.empty(T, dtype=np.int64)
    X[-1] = np.argmax(V)
    for t in range(T - 1, 0, -1):
        X[t - 1] = P[t, X[t]]
    return X.astype(np.float64)


def viterbi_c(emission, transition):
    emission = np.asarray(emission)
    T, K = emission.shape
    trans = np.ascontiguousarray(transition, dtype=np.float64)
    out = np.empty(T, dtype=np.int32)
    if T == 0:
        return out.astype(np.float64)
    lib = _lib()
    if emission.dtype == np.float32:
        em = np.ascontiguousarray(emission)
        rc = lib.oracle_viterbi_f32(em.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                    trans.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                    T, K, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    else:
        em = np.ascontiguousarray(emission, dtype=np.float64)
        rc = lib.oracle_viterbi(em.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                trans.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                T, K, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise RuntimeError('oracle_viterbi failed rc=%d' % rc)
    return out.astype(np.float64)


def viterbi_decoding(emission, transition):
    """Drop-in for the reference call signature used on the hot path."""
    return viterbi_c(emission, transition)
