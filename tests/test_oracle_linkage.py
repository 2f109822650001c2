"""The linkage oracle (oracle/linkage.py) against scipy.cluster.hierarchy.linkage itself, bit for bit."""
import gzip
import json
import os

import numpy as np
import pytest
from scipy.cluster.hierarchy import linkage
from scipy.spatial.distance import squareform

from oracle import linkage as L
from conftest import GOLDEN


def matrices():
    rng = np.random.default_rng(2)
    out = []
    for n in (2, 3, 5, 13, 40, 101):
        M = rng.random((n, n))
        out.append(('random', np.minimum(M, M.T)))
        out.append(('ties', np.round(np.minimum(M, M.T) * 6) / 6))                 # many equal distances
        F = (rng.integers(1, 100001, (n, n)) / 1e4).astype('<f4')                  # float32 values in [1e-4, 10] like the real matrices
        out.append(('float32', np.minimum(F, F.T).astype('<f4') / np.float32(10.0)))
    out.append(('constant', np.full((17, 17), 0.25)))
    return out


@pytest.mark.parametrize('method', ['ward', 'single', 'complete', 'average', 'weighted'])
def test_oracle_linkage_equals_scipy(method):
    for name, M in matrices():
        M = M.copy()
        np.fill_diagonal(M, 0)
        y = squareform(M, checks=False)
        want = linkage(y, method=method)
        got = L.linkage(y, method=method)
        assert got.shape == want.shape and np.array_equal(got, want), (method, name, len(M))


def test_oracle_linkage_on_the_bundled_distance_matrices():
    D = np.load(os.path.join(GOLDEN, 'dp_golden.npz'))
    n_checked = 0
    for key in D.files:
        M = D[key]
        if M.ndim == 2 and M.shape[0] == M.shape[1] and M.shape[0] >= 3 and M.dtype == np.dtype('<f4'):
            M = M / np.float32(10.0)                                                 # tg_tvr.py:154 (float32, in place upstream)
            y = squareform(M, checks=False)
            for method in ('ward', 'single'):
                assert np.array_equal(L.linkage(y, method=method), linkage(y, method=method)), (key, method)
            n_checked += 1
    assert n_checked >= 2
