"""GPU: tg_linkage.linkage / linkage_from_matrix against scipy.cluster.hierarchy.linkage itself -- the merge lists must be
bit-identical (same algorithms, same float64 evaluation order), so fcluster gives the same alleles."""
import os

import numpy as np
import pytest
from scipy.cluster.hierarchy import fcluster, linkage
from scipy.spatial.distance import squareform

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

METHODS = ['ward', 'single', 'complete', 'average', 'weighted']


def cases():
    rng = np.random.default_rng(11)
    out = []
    for n in (2, 3, 7, 33, 100, 257, 1025, 1500):
        M = rng.random((n, n))
        out.append((f'random{n}', np.minimum(M, M.T)))
    for n in (5, 64, 300):
        M = rng.random((n, n))
        out.append((f'ties{n}', np.round(np.minimum(M, M.T) * 5) / 5))
        F = (rng.integers(1, 100001, (n, n)) / 1e4).astype('<f4')
        out.append((f'float32_{n}', (np.minimum(F, F.T) / np.float32(10.0)).astype(np.float64)))
    out.append(('constant', np.full((40, 40), 0.5)))
    blocks = np.kron(np.eye(4), np.ones((25, 25))) * -0.8 + 0.9 + rng.random((100, 100)) * 0.05       # four tight clusters
    out.append(('blocks', np.minimum(blocks, blocks.T)))
    return out


@pytest.mark.parametrize('method', METHODS)
def test_linkage_equals_scipy(method):
    from telogator2_b200 import tg_linkage
    for name, M in cases():
        M = M.copy()
        np.fill_diagonal(M, 0)
        y = squareform(M, checks=False)
        want = linkage(y, method=method)
        got = tg_linkage.linkage(y, method=method)
        assert got.dtype == np.float64 and got.shape == want.shape, (method, name)
        assert np.array_equal(got, want), (method, name)


def test_linkage_from_the_float32_matrix_and_flat_clusters():
    """The call sites of the reference: float32 matrix / 10 -> squareform -> linkage -> fcluster (tg_tvr.py:154-186, 397-398)."""
    import torch
    from telogator2_b200 import tg_linkage
    D = np.load(os.path.join(GOLDEN, 'dp_golden.npz'))
    n_checked = 0
    for key in D.files:
        M = D[key]
        if not (M.ndim == 2 and M.shape[0] == M.shape[1] and M.shape[0] >= 3 and M.dtype == np.dtype('<f4')):
            continue
        ref = M.copy()
        ref /= 10.0                                                    # the reference's in-place float32 division
        for method, cut in (('ward', 0.2), ('single', 0.05), ('complete', 0.1)):
            want = linkage(squareform(ref, checks=False), method=method)
            got_h = tg_linkage.linkage(squareform(ref, checks=False), method=method)
            got_m = tg_linkage.linkage_from_matrix(M, method=method, divide_by=10.0)
            got_d = tg_linkage.linkage_from_matrix(torch.from_numpy(M).cuda(), method=method, divide_by=10.0)
            assert np.array_equal(got_h, want) and np.array_equal(got_m, want) and np.array_equal(got_d, want), (key, method)
            assert np.array_equal(fcluster(got_d, cut, 'distance'), fcluster(want, cut, 'distance'))
        n_checked += 1
    assert n_checked >= 2


@pytest.mark.parametrize('ctas', ['2', '7', '148'])
def test_multi_cta_runs_equal_scipy(monkeypatch, ctas):
    """Long rows are walked by several CTAs with a grid barrier between the dependent steps (tg_linkage_run_multi); forced here on
    small matrices, every method."""
    from telogator2_b200 import tg_linkage
    monkeypatch.setenv('TG_LINKAGE_CTAS', ctas)
    for name, M in cases():
        if len(M) < 3:
            continue
        M = M.copy()
        np.fill_diagonal(M, 0)
        y = squareform(M, checks=False)
        for method in METHODS:
            assert np.array_equal(tg_linkage.linkage(y, method=method), linkage(y, method=method)), (ctas, method, name)


def test_large_matrix_takes_the_multi_cta_path():
    from telogator2_b200 import tg_linkage
    n = 9000
    assert tg_linkage._multi_ctas(n) > 1
    rng = np.random.default_rng(4)
    M = (rng.integers(1, 100001, (n, n), dtype=np.int32) / 1e4).astype('<f4')
    M = np.minimum(M, M.T)
    np.fill_diagonal(M, 0)
    ref = M / np.float32(10.0)
    for method in ('ward', 'single'):
        assert np.array_equal(tg_linkage.linkage_from_matrix(M, method, divide_by=10.0), linkage(squareform(ref, checks=False), method))


def test_unsupported_calls_fail_loudly():
    from telogator2_b200 import tg_linkage
    y = squareform(np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0.0]]))
    for method in ('centroid', 'median', 'nonsense'):
        with pytest.raises(NotImplementedError):
            tg_linkage.linkage(y, method=method)
    with pytest.raises(NotImplementedError):
        tg_linkage.linkage(np.zeros((4, 2)), method='ward')
    with pytest.raises(ValueError):
        tg_linkage.linkage(np.zeros(4), method='ward')
    with pytest.raises(ValueError):
        tg_linkage.linkage(np.array([1.0, np.nan, 2.0]), method='ward')
