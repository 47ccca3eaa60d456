"""Oracle: literal NumPy restatement of the reference skeleton-graph construction.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows net/utils/graph.py of the reference
step by step -- matrix powers for the hop distances (:396-420), the nested i/j loops for the
partitions (:165-187), the in-place neighbour expansion (:508-518) -- so that it shares no
code path with the vectorised product implementation in ist-gcn_b200/net/utils/graph.py.
Only the layouts/strategies that run in the reference are restated.
"""
import copy

import numpy as np

_OPENPOSE = [(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5),
             (8, 2), (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14)]
_OPENPOSE_SYM = [(14, 15), (16, 17), (2, 5), (3, 6), (4, 7), (8, 11), (9, 12), (10, 13)]
_NTU = [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9),
        (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18),
        (20, 19), (22, 23), (23, 8), (24, 25), (25, 12)]
_NTU_SYM = [(23, 25), (24, 22), (11, 7), (10, 6), (9, 5), (8, 12), (16, 20), (17, 13), (18, 14),
            (19, 15)]
_NTU_HALF = [(1, 2), (2, 13), (3, 13), (4, 3), (5, 13), (6, 5), (7, 6), (8, 7), (9, 1), (10, 9),
             (11, 10), (12, 11), (14, 15), (15, 8)]


def layout(name):
    """graph.py:47-143 -> (num_node, edge list, centre, mirror pairs)."""
    if name in ('openpose', 'openpose_sym'):
        n, links, center, sym = 18, list(_OPENPOSE), 1, list(_OPENPOSE_SYM)
    elif name == 'ntu-rgb+d':
        n, links, center, sym = 25, [(a - 1, b - 1) for a, b in _NTU], 20, []
    elif name == 'ntu-rgb+d_sym':
        n, links, center = 25, [(a - 1, b - 1) for a, b in _NTU], 20
        sym = [(a - 1, b - 1) for a, b in _NTU_SYM]
    elif name == 'ntu-rgb+d_half':
        n, links, center, sym = 15, [(a - 1, b - 1) for a, b in _NTU_HALF], 12, []
    else:
        raise ValueError("Do Not Exist This Layout.")
    return n, [(i, i) for i in range(n)] + links, center, sym


def hop_matrix(adj, top):
    """graph.py:396-420: d = smallest power with a positive entry, inf if none up to ``top``."""
    n = adj.shape[0]
    powers = np.stack([np.linalg.matrix_power(adj, d) for d in range(top + 1)]) > 0
    dist = np.zeros((n, n)) + np.inf
    for d in range(top, -1, -1):
        dist[powers[d]] = d
    return dist


def column_normalise(adj):
    """graph.py:453-461."""
    deg = np.sum(adj, 0)
    n = adj.shape[0]
    dn = np.zeros((n, n))
    for i in range(n):
        if deg[i] > 0:
            dn[i, i] = deg[i] ** (-1)
    return np.dot(adj, dn)


def norm_at(hop, dist):
    """graph.py:498-505."""
    ind = np.zeros(dist.shape)
    for h in (0, hop):
        ind[dist == h] = 1
    return column_normalise(ind)


def expand(adjacency, stack, norm, n, kernel_size):
    """graph.py:508-525 (get_A + add_one_distance), scan order preserved."""
    res = copy.deepcopy(stack)
    for part in range(1, kernel_size):
        for i in range(n):
            for j in range(n):
                if res[part][j, i] != 0:
                    res[part][j, i] = norm[j, i]
                    for k in range(n):
                        if adjacency[j][k] == 1 and res[1][k, i] == 0 and k != i:
                            res[part][k, i] = norm[k, i]
    return res


def mirror_partition(stack, norm, n, pairs):
    """graph.py:528-536."""
    extra = np.zeros((n, n))
    for i, j in pairs:
        extra[i, j] = norm[i, j]
    return np.append(stack, np.expand_dims(extra, 0), axis=0)


def build(layout_name, strategy, max_hop=3, dilation=1, kernel_size=3):
    """Graph(layout, strategy).A[, A2, A3] -- graph.py:27-42,145-361.  Returns a dict."""
    n, edge, center, sym = layout(layout_name)
    adj = np.zeros((n, n))
    for i, j in edge:
        adj[j, i] = 1
        adj[i, j] = 1
    adj_sym = copy.deepcopy(adj)
    for i, j in sym:
        adj_sym[j, i] = 1
        adj_sym[i, j] = 1
    dist = hop_matrix(adj, n)
    dist_sym = hop_matrix(adj_sym, n)
    n1, n2, n3 = norm_at(1, dist_sym), norm_at(2, dist), norm_at(3, dist)
    hops = range(0, 2, dilation)
    out = {}
    if strategy == 'uniform':
        out['A'] = n1[None]
        return out
    if strategy == 'distance':
        A = np.zeros((len(hops), n, n))
        for idx, hop in enumerate(hops):
            A[idx][dist == hop] = n1[dist == hop]
        out['A'] = A
        return out
    if strategy not in ('spatial', 'spatial_half', 'spatial_3', 'spatial_sym', 'spatial_3_sym'):
        raise ValueError("Do Not Exist This Strategy")
    parts = []
    for hop in hops:
        root, close, further = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        for i in range(n):
            for j in range(n):
                if dist[j, i] == hop:
                    if dist[j, center] == dist[i, center]:
                        root[j, i] = n1[j, i]
                    elif dist[j, center] > dist[i, center]:
                        close[j, i] = n1[j, i]
                    else:
                        further[j, i] = n1[j, i]
        if hop == 0:
            parts.append(root)
        else:
            parts.append(root + close)
            parts.append(further)
    A = np.stack(parts)
    if strategy in ('spatial', 'spatial_half'):
        out['A'] = A
    elif strategy == 'spatial_sym':
        out['A'] = mirror_partition(A, n2, n, sym)
    else:
        A2 = expand(adj, A, n2, n, kernel_size)
        A3 = expand(adj, A2, n3, n, kernel_size)
        if strategy == 'spatial_3_sym':
            A = mirror_partition(A, n1, n, sym)
            zero = np.zeros((1, n, n))
            A2 = np.append(A2, zero, axis=0)
            A3 = np.append(A3, zero, axis=0)
        out['A'], out['A2'], out['A3'] = A, A2, A3
    return out
