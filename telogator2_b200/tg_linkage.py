"""scipy.cluster.hierarchy.linkage on the GPU for the calls the reference makes (tg_tvr.py:166-167, 397-398).

Drop-in for `linkage(y, method)` with a condensed distance array and method ward / single / complete / average / weighted:
the returned (n - 1, 4) float64 array is bit-identical to scipy's, so `fcluster`, `dendrogram` and everything downstream
behave as before (they stay scipy's own: O(n) on the merge list).  Other methods, observation matrices and
optimal_ordering are not on the reference's path and raise NotImplementedError (the overlay passes those to scipy).

`linkage_from_matrix` clusters straight from a square float32 matrix -- the device matrix of the distance kernels, or the
host matrix get_dist_matrix_parallel returned -- without the squareform / float64 copy round trip.
"""
import numpy as np

from . import _lib

METHODS = ('ward', 'single', 'complete', 'average', 'weighted')


def _label(Z, n):
    from .tg_kmer import _hostlists
    H = _hostlists()
    if H is not None and hasattr(H, 'linkage_label'):
        H.linkage_label(Z, n)
        return
    parent = list(range(2 * n - 1))               # scipy _hierarchy.label, pure Python (the C helper has not been built)
    size = [1] * (2 * n - 1)
    nxt = n
    for i in range(n - 1):
        r = []
        for s in (0, 1):
            x = int(Z[i, s])
            p = x
            while parent[x] != x:
                x = parent[x]
            while parent[p] != x:
                p, parent[p] = parent[p], x
            r.append(x)
        Z[i, 0], Z[i, 1] = min(r), max(r)
        parent[r[0]] = parent[r[1]] = nxt
        size[nxt] = size[r[0]] + size[r[1]]
        Z[i, 3] = size[nxt]
        nxt += 1


MULTI_CTA_MIN_N = 5120          # below this one CTA is as fast (a grid barrier costs as much as walking a short row: 4096 -> 56 vs 67 ms)


def _multi_ctas(n):
    """CTAs for the row walks: 1 for short rows, else about one per 1024 columns (TG_LINKAGE_CTAS overrides)."""
    import os
    env = os.environ.get('TG_LINKAGE_CTAS')
    torch = _lib.require_cuda()
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    if env is not None:
        return max(1, min(int(env), sms))
    if n < MULTI_CTA_MIN_N:
        return 1
    return max(2, min(sms, (n + 1023) // 1024))


def _run(d_square, n, method):
    torch = _lib.require_cuda()
    L = _lib.load()
    if method not in METHODS:
        raise NotImplementedError(f"linkage method '{method}' is not on the reference's path (supported: {', '.join(METHODS)})")
    d_size = torch.ones(n, dtype=torch.int32, device='cuda')
    d_near = torch.empty(n, dtype=torch.float64, device='cuda')
    d_Z = torch.empty((n - 1, 4), dtype=torch.float64, device='cuda')
    ctas = _multi_ctas(n)
    if ctas > 1:
        # long rows: several CTAs walk each row, a grid barrier separates the dependent steps
        d_work = torch.empty(L.tg_linkage_multi_work_bytes(n, ctas), dtype=torch.uint8, device='cuda')
        _lib.check(L.tg_linkage_run_multi(_lib.dptr(d_square), n, method.encode(), ctas, _lib.dptr(d_size), _lib.dptr(d_near),
                                          _lib.dptr(d_Z), _lib.dptr(d_work), _lib.stream_ptr(torch)))
    else:
        d_chain = torch.zeros(n, dtype=torch.int32, device='cuda')
        _lib.check(L.tg_linkage_run(_lib.dptr(d_square), n, method.encode(), _lib.dptr(d_size), _lib.dptr(d_chain), _lib.dptr(d_near),
                                    _lib.dptr(d_Z), _lib.stream_ptr(torch)))
    Z = d_Z.cpu().numpy()
    Z = np.ascontiguousarray(Z[np.argsort(Z[:, 2], kind='mergesort')])      # _hierarchy: "Sort Z by cluster distances" (stable)
    _label(Z, n)
    return Z


def linkage(y, method='single', metric='euclidean', optimal_ordering=False):
    """scipy.cluster.hierarchy.linkage(y, method) for a condensed distance array."""
    if optimal_ordering:
        raise NotImplementedError('optimal_ordering is not on the reference\'s path')
    y = np.asarray(y)
    if y.ndim != 1:
        raise NotImplementedError('linkage of an observation matrix is not on the reference\'s path (pass the condensed distances)')
    y = np.ascontiguousarray(y, dtype=np.float64)                          # scipy: _convert_to_double
    m = len(y)
    n = int((1 + np.sqrt(1 + 8 * m)) // 2) if m else 1
    if n * (n - 1) // 2 != m or m == 0:
        raise ValueError('Incompatible vector size. It must be a binomial coefficient n choose 2 for some integer n >= 2.')
    if not np.all(np.isfinite(y)):
        raise ValueError('The condensed distance matrix must contain only finite values.')
    torch = _lib.require_cuda()
    L = _lib.load()
    d_y = torch.from_numpy(y).cuda()
    d_sq = torch.empty((n, n), dtype=torch.float64, device='cuda')
    _lib.check(L.tg_linkage_expand(_lib.dptr(d_y), n, _lib.dptr(d_sq), _lib.stream_ptr(torch)))
    return _run(d_sq, n, method)


def linkage_from_matrix(dist_matrix, method='ward', divide_by=0.0):
    """linkage(squareform(dist_matrix / divide_by), method) from a square float32 matrix (torch CUDA tensor or ndarray);
    divide_by != 0 applies the reference's float32 `dist_matrix /= MAX_SEQ_DIST` (tg_tvr.py:154) on the way."""
    torch = _lib.require_cuda()
    L = _lib.load()
    M = dist_matrix if hasattr(dist_matrix, 'is_cuda') else torch.from_numpy(np.ascontiguousarray(dist_matrix, dtype=np.float32))
    if M.dim() != 2 or M.shape[0] != M.shape[1] or M.dtype != torch.float32:
        raise ValueError('linkage_from_matrix: a square float32 matrix is required')
    n = int(M.shape[0])
    if n < 2:
        raise ValueError('at least two observations are required')
    M = M.cuda().contiguous()
    d_sq = torch.empty((n, n), dtype=torch.float64, device='cuda')
    _lib.check(L.tg_linkage_widen(_lib.dptr(M), n, float(divide_by), _lib.dptr(d_sq), _lib.stream_ptr(torch)))
    return _run(d_sq, n, method)
