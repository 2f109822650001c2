"""CPU oracle of the hierarchical clustering that consumes the distance matrix -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls scipy: `linkage(squareform(dist_matrix), method='ward')` (tg_tvr.py:166-167) and `linkage(d_arr,
method=linkage_method)` with 'single' / 'ward' / 'complete' (tg_tvr.py:397-398, telogator2.py call sites), then `fcluster(Z,
tree_cut, 'distance')`.  scipy IS installed (here and on the GPU boxes), so this restatement is PINNED: tests compare it, and the
CUDA path, bit for bit with scipy.cluster.hierarchy.linkage itself.

Restated from scipy/cluster/_hierarchy.pyx (scipy >= 1.0): `nn_chain` (ward, complete, average, weighted), `mst_single_linkage`
(single), the stable sort of the merges by distance and `label` (union-find relabelling).  Pure Python loops: small inputs only.
"""
import math

import numpy as np


def condensed_index(n, i, j):
    if i < j:
        return n * i - (i * (i + 1) // 2) + (j - i - 1)
    return n * j - (j * (j + 1) // 2) + (i - j - 1)


def _ward(d_xi, d_yi, d_xy, size_x, size_y, size_i):
    t = 1.0 / (size_x + size_y + size_i)
    return math.sqrt((size_i + size_x) * t * d_xi * d_xi + (size_i + size_y) * t * d_yi * d_yi - size_i * t * d_xy * d_xy)


def _complete(d_xi, d_yi, d_xy, size_x, size_y, size_i):
    return max(d_xi, d_yi)


def _average(d_xi, d_yi, d_xy, size_x, size_y, size_i):
    return (size_x * d_xi + size_y * d_yi) / (size_x + size_y)


def _weighted(d_xi, d_yi, d_xy, size_x, size_y, size_i):
    return 0.5 * (d_xi + d_yi)


NEW_DIST = {'ward': _ward, 'complete': _complete, 'average': _average, 'weighted': _weighted}


def label(Z, n):
    """_hierarchy.label: cluster ids of the merged pairs (smaller root first) and cluster sizes, in place."""
    parent = list(range(2 * n - 1))
    size = [1] * (2 * n - 1)
    nxt = n

    def find(x):
        p = x
        while parent[x] != x:
            x = parent[x]
        while parent[p] != x:
            p, parent[p] = parent[p], x
        return x
    for i in range(n - 1):
        xr, yr = find(int(Z[i, 0])), find(int(Z[i, 1]))
        Z[i, 0], Z[i, 1] = (xr, yr) if xr < yr else (yr, xr)
        parent[xr] = parent[yr] = nxt
        size[nxt] = size[xr] + size[yr]
        Z[i, 3] = size[nxt]
        nxt += 1


def nn_chain(dists, n, method):
    new_dist = NEW_DIST[method]
    D = np.array(dists, dtype=np.float64)
    size = [1] * n
    Z = np.zeros((n - 1, 4))
    chain = [0] * n
    chain_length = 0
    for k in range(n - 1):
        if chain_length == 0:
            chain_length = 1
            for i in range(n):
                if size[i] > 0:
                    chain[0] = i
                    break
        while True:
            x = chain[chain_length - 1]
            if chain_length > 1:                         # prefer the previous element of the chain on ties
                y = chain[chain_length - 2]
                current_min = D[condensed_index(n, x, y)]
            else:
                current_min = float('inf')
            for i in range(n):
                if size[i] == 0 or x == i:
                    continue
                dist = D[condensed_index(n, x, i)]
                if dist < current_min:
                    current_min = dist
                    y = i
            if chain_length > 1 and y == chain[chain_length - 2]:
                break
            chain[chain_length] = y
            chain_length += 1
        chain_length -= 2
        if x > y:
            x, y = y, x
        nx, ny = size[x], size[y]
        Z[k] = [x, y, current_min, nx + ny]
        size[x] = 0
        size[y] = nx + ny
        for i in range(n):
            ni = size[i]
            if ni == 0 or i == y:
                continue
            D[condensed_index(n, i, y)] = new_dist(D[condensed_index(n, i, x)], D[condensed_index(n, i, y)], current_min, nx, ny, ni)
    Z = Z[np.argsort(Z[:, 2], kind='mergesort')]
    label(Z, n)
    return Z


def mst_single_linkage(dists, n):
    dists = np.asarray(dists, dtype=np.float64)
    Z = np.zeros((n - 1, 4))
    merged = [0] * n
    D = [float('inf')] * n
    x = y = 0
    for k in range(n - 1):
        current_min = float('inf')
        merged[x] = 1
        for i in range(n):
            if merged[i] == 1:
                continue
            dist = dists[condensed_index(n, x, i)]
            if D[i] > dist:
                D[i] = dist
            if D[i] < current_min:
                y = i
                current_min = D[i]
        Z[k, 0], Z[k, 1], Z[k, 2] = x, y, current_min
        x = y
    Z = Z[np.argsort(Z[:, 2], kind='mergesort')]
    label(Z, n)
    return Z


def linkage(y, method='single'):
    """scipy.cluster.hierarchy.linkage(y, method) for a condensed distance array and the methods the reference uses."""
    y = np.asarray(y, dtype=np.float64)
    n = int((1 + math.isqrt(1 + 8 * len(y))) // 2)
    if n * (n - 1) // 2 != len(y):
        raise ValueError('not a condensed distance matrix')
    if method == 'single':
        return mst_single_linkage(y, n)
    return nn_chain(y, n, method)
