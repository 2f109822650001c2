"""The reference's source/tg_tvr.py; cluster_tvrs / cluster_consensus_tvrs reach the GPU through the names they import from
source.tg_align.  quick_get_tvrtel_lens (tg_tvr.py:559-594) is bound to the batched colour-vector kernels, and `linkage`
(scipy, tg_tvr.py:167, 398) to the device implementation for the methods it covers (same merge list, bit for bit)."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

_scipy_linkage = linkage  # noqa: F821  (scipy.cluster.hierarchy.linkage, imported by the reference module)

from telogator2_b200.tg_tvr import quick_get_tvrtel_lens  # noqa: E402,F401


def linkage(y, method='single', metric='euclidean', optimal_ordering=False):
    from telogator2_b200 import tg_linkage
    import numpy as _np
    if method in tg_linkage.METHODS and not optimal_ordering and _np.ndim(y) == 1 and len(y) >= 3:
        return tg_linkage.linkage(y, method=method)
    return _scipy_linkage(y, method=method, metric=metric, optimal_ordering=optimal_ordering)
