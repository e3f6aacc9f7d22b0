"""Pool-adjacent-violators isotonic regression (oracle; test infrastructure only).

Restates circuitmap/optimise/pava.py:9-88 for the only case the hot path uses
(gamma = 1, unit weights; caviar.py:128,220 consume element [-1] of the fit).
"""
import numpy as np


def make_pava_pools(y):
    """pava.py:9-61 with gamma=1 (lg=0): stack of pools (v=sum, w=weight, l=length).

    A new singleton pool is pushed per element; while the previous pool's mean is
    STRICTLY greater than the last pool's mean (pava.py:42) the two are merged, the
    sums accumulating in merge order (pava.py:49-51).
    """
    y = np.asarray(y, dtype=np.float64)
    T = len(y)
    v = np.zeros(T)
    w = np.zeros(T)
    l = np.zeros(T, dtype=np.int32)
    v[0], w[0], l[0] = y[0], 1.0, 1
    i = 0
    for t in range(1, T):
        i += 1
        v[i], w[i], l[i] = y[t], 1.0, 1
        while i > 0 and (v[i - 1] / w[i - 1]) > (v[i] / w[i]):
            i -= 1
            v[i] = v[i] + v[i + 1]
            w[i] = w[i] + w[i + 1]
            l[i] = l[i] + l[i + 1]
            v[i + 1] = w[i + 1] = 0.0
            l[i + 1] = 0
    return v, w, l


def isotonic_regression(y):
    """pava.py:63-88: expand the pools back to a length-T non-decreasing fit."""
    y = np.asarray(y, dtype=np.float64)
    v, w, l = make_pava_pools(y)
    out = np.zeros_like(y)
    t = 0
    for i in range(len(v)):
        if l[i] > 0:
            out[t:t + l[i]] = v[i] / w[i]
            t += l[i]
    return out


def pava_last(y):
    """isotonic_regression(y)[-1] -- mean of the last pool."""
    v, w, l = make_pava_pools(y)
    i = np.nonzero(l)[0][-1]
    return v[i] / w[i]
