"""Skeleton graph: joint layouts, hop distances and the partitioned adjacency stacks.

Drop-in for the reference ``net.utils.graph.Graph`` (reference: net/utils/graph.py:5-361,
helpers :364-536).  Runs once per ``Model.__init__`` on the host in float64 NumPy; the
result is cast to fp32 and registered as the ``A`` / ``A2`` / ``A3`` buffers, so it must be
*bit-exact* with the reference (tests/test_graph.py pins it against the golden SHA table of
SURVEY.md App. A, the committed tests/golden/graphs.npz and -- when /root/reference is
mounted -- the live reference).

This is a re-derivation, not a transcription: layouts live in one table, hop distances come
from a breadth-first frontier expansion, and the three-way root/centripetal/centrifugal split
is done with boolean masks.  The places where the reference's semantics are *order dependent*
(the in-place neighbour expansion that builds A2/A3, graph.py:508-518) are restated step by
step because the scan order is part of the result.

Semantics worth knowing (all verified against the reference):
  * partition membership uses the FULL shortest-path distance on skeleton edges
    (graph.py:416-420,445), not upstream ST-GCN's distance capped at ``max_hop``;
  * the values come from ``normalize_adjacency1`` built on skeleton + mirror edges
    (graph.py:151), so with a non-empty mirror list columns no longer sum to one;
  * the valid hops are always {0, 1} (graph.py:146); ``max_hop`` only feeds ``hop_dis23``;
  * ``spatial_sym`` fills the 4th partition from the distance-2 normalisation, one direction
    only (graph.py:323, :530-531); ``spatial_3_sym`` from ``normalize_adjacency1`` (:350) and
    appends an all-zero 4th partition to A2 / A3 (:353-356).
"""
import numpy as np

# name -> (num_node, one_based?, neighbour links, centre joint (0-based), mirror pairs or None)
# ``None`` mirrors the reference layouts that never define ``spatial_symmetric``
# (graph.py:57-70, 100-115, 129-140): constructing a Graph with them raises AttributeError.
_OPENPOSE_LINKS = ((4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8),
                   (11, 5), (8, 2), (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14))
_OPENPOSE_MIRROR = ((14, 15), (16, 17), (2, 5), (3, 6), (4, 7), (8, 11), (9, 12), (10, 13))
_NTU_LINKS_1B = ((1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21),
                 (10, 9), (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1),
                 (18, 17), (19, 18), (20, 19), (22, 23), (23, 8), (24, 25), (25, 12))
_NTU_MIRROR_1B = ((23, 25), (24, 22), (11, 7), (10, 6), (9, 5), (8, 12), (16, 20), (17, 13),
                  (18, 14), (19, 15))
_NTU_HALF_LINKS_1B = ((1, 2), (2, 13), (3, 13), (4, 3), (5, 13), (6, 5), (7, 6), (8, 7), (9, 1),
                      (10, 9), (11, 10), (12, 11), (14, 15), (15, 8))
_NTU_EDGE_LINKS_1B = ((1, 2), (3, 2), (4, 3), (5, 2), (6, 5), (7, 6), (8, 7), (9, 2), (10, 9),
                      (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1),
                      (18, 17), (19, 18), (20, 19), (21, 22), (22, 8), (23, 24), (24, 12))


def _shift(pairs):
    return [(i - 1, j - 1) for (i, j) in pairs]


def _layout_table():
    hub18 = tuple((18, j) for j in range(18))
    hub26 = tuple((26, j) for j in range(1, 26))
    return {
        'openpose': (18, list(_OPENPOSE_LINKS), 1, list(_OPENPOSE_MIRROR)),
        'openpose_sym': (18, list(_OPENPOSE_LINKS), 1, list(_OPENPOSE_MIRROR)),
        'openpose_gravity': (19, list(_OPENPOSE_LINKS + hub18), 1, None),
        'ntu-rgb+d': (25, _shift(_NTU_LINKS_1B), 20, []),
        'ntu-rgb+d_sym': (25, _shift(_NTU_LINKS_1B), 20, _shift(_NTU_MIRROR_1B)),
        'ntu-rgb+d_half': (15, _shift(_NTU_HALF_LINKS_1B), 12, []),
        'ntu-rgb+d_gravity': (26, _shift(_NTU_LINKS_1B + hub26), 20, None),
        'ntu_edge': (24, _shift(_NTU_EDGE_LINKS_1B), 2, None),
    }


class Graph():
    """Graph(layout, strategy, max_hop=3, dilation=1, kernel_size=3)  -- graph.py:27-42.

    Attributes (same names as the reference): ``num_node``, ``edge``, ``center``,
    ``spatial_symmetric``, ``adjacency_matrix``, ``hop_dis``, ``hop_dis_sym``, ``hop_dis23``,
    ``A`` (K,V,V) float64 and, for the ``spatial_3*`` strategies, ``A2`` / ``A3``.
    """

    def __init__(self, layout='openpose', strategy='uniform', max_hop=3, dilation=1,
                 kernel_size=3):
        self.max_hop = max_hop
        self.dilation = dilation
        self.kernel_size = kernel_size
        self.get_edge(layout)
        (self.adjacency_matrix, self.hop_dis, self.hop_dis_sym,
         self.hop_dis23) = get_hop_distance(self.num_node, self.edge, self.spatial_symmetric,
                                            max_hop=max_hop)
        self.get_adjacency(strategy)

    def get_edge(self, layout):
        """graph.py:47-143."""
        table = _layout_table()
        if layout not in table:
            raise ValueError("Do Not Exist This Layout.")
        num_node, links, center, mirror = table[layout]
        self.num_node = num_node
        self.edge = [(i, i) for i in range(num_node)] + links
        self.center = center
        if mirror is not None:
            self.spatial_symmetric = mirror
            if layout == 'ntu-rgb+d_sym':
                self.spatial_symmetric1 = list(_NTU_MIRROR_1B)

    # -- partitions -----------------------------------------------------------------------
    def _three_way(self, norm, limit=None):
        """Self / (same-or-farther-from-centre source) / closer-source split of the hop-0 and
        hop-1 entries: graph.py:165-187 (and the identical loops of the other strategies).
        Entry [j, i] is classified by comparing hop_dis[j, centre] with hop_dis[i, centre].
        ``limit`` restricts both indices to < limit (the *_gravity strategies, :219-220)."""
        n = self.num_node
        d_center = self.hop_dis[:, self.center]
        src = d_center[:, None]          # distance of row joint j to the centre
        dst = d_center[None, :]          # distance of column joint i to the centre
        inside = np.ones((n, n), dtype=bool)
        if limit is not None:
            inside[limit:, :] = False
            inside[:, limit:] = False
        stack = []
        for hop in range(0, 2, self.dilation):
            at_hop = (self.hop_dis == hop) & inside
            root = np.where(at_hop & (src == dst), norm, 0.0)
            close = np.where(at_hop & (src > dst), norm, 0.0)
            further = np.where(at_hop & (src < dst), norm, 0.0)
            if hop == 0:
                stack.append(root)
            else:
                stack.append(root + close)
                stack.append(further)
        return np.stack(stack)

    def get_adjacency(self, strategy):
        """graph.py:145-361."""
        n = self.num_node
        valid_hop = range(0, 2, self.dilation)
        norm1 = get_norm(1, self.hop_dis_sym, n, self.dilation)
        norm2 = get_norm(2, self.hop_dis, n, self.dilation)
        norm3 = get_norm(3, self.hop_dis, n, self.dilation)
        if strategy == 'uniform':
            self.A = norm1[None].copy()
        elif strategy == 'distance':
            A = np.zeros((len(valid_hop), n, n))
            for idx, hop in enumerate(valid_hop):
                sel = self.hop_dis == hop
                A[idx][sel] = norm1[sel]
            self.A = A
        elif strategy in ('spatial', 'spatial_half'):
            self.A = self._three_way(norm1)
        elif strategy in ('openpose_gravity', 'ntu-rgb+d_gravity'):
            hub = 18 if strategy == 'openpose_gravity' else 25
            A = self._three_way(norm1, limit=n - 1)
            gravity = np.zeros((n, n))
            gravity[hub, :] = norm1[hub, :]      # IndexError when the layout has no hub joint,
            gravity[:, hub] = norm1[:, hub]      # exactly like graph.py:239-241 / :270-272
            self.A = np.concatenate([A, gravity[None]], axis=0)
        elif strategy == 'spatial_3':
            A = self._three_way(norm1)
            A2 = get_A(A, norm2, self.adjacency_matrix, n, self.kernel_size)
            A3 = get_A(A2, norm3, self.adjacency_matrix, n, self.kernel_size)
            self.A, self.A2, self.A3 = A, A2, A3
        elif strategy == 'spatial_sym':
            A = self._three_way(norm1)
            self.A = every_symmetric(A, norm2, n, self.spatial_symmetric)
        elif strategy == 'spatial_3_sym':
            A = self._three_way(norm1)
            A2 = get_A(A, norm2, self.adjacency_matrix, n, self.kernel_size)
            A3 = get_A(A2, norm3, self.adjacency_matrix, n, self.kernel_size)
            A = every_symmetric(A, norm1, n, self.spatial_symmetric)
            empty = np.zeros((1, n, n))
            self.A = A
            self.A2 = np.concatenate([A2, empty], axis=0)
            self.A3 = np.concatenate([A3, empty], axis=0)
        else:
            raise ValueError("Do Not Exist This Strategy")


def _bfs_distance(adj, cap):
    """Smallest d <= cap with a walk of length d (self loops allowed) between each pair;
    inf where none exists.  Equals the reference's ``matrix_power(A, d) > 0`` sweep
    (graph.py:396-420) because every joint carries a self loop."""
    n = adj.shape[0]
    link = adj > 0
    dist = np.full((n, n), np.inf)
    reach = np.eye(n, dtype=bool)
    dist[reach] = 0
    for d in range(1, cap + 1):
        grown = (reach.astype(np.int64) @ link.astype(np.int64)) > 0
        dist[grown & ~reach] = d
        if np.array_equal(grown | reach, reach):
            break
        reach = grown | reach
    return dist


def get_hop_distance(num_node, edge, spatial_symmetric, max_hop=1):
    """graph.py:364-445 -> (adjacency_matrix, hop_dis_all, hop_dis_sym, hop_dis23)."""
    adj = np.zeros((num_node, num_node))
    for i, j in edge:
        adj[j, i] = 1
        adj[i, j] = 1
    with_mirror = adj.copy()
    for i, j in spatial_symmetric:
        with_mirror[j, i] = 1
        with_mirror[i, j] = 1
    hop_dis_sym = _bfs_distance(with_mirror, num_node)
    hop_dis23 = _bfs_distance(adj, max_hop)
    hop_dis_all = _bfs_distance(adj, num_node)
    return adj, hop_dis_all, hop_dis_sym, hop_dis23


def normalize_digraph(A):
    """Column normalisation A @ diag(1/colsum) -- graph.py:453-461.  The product has a single
    non-zero term per entry, so scaling the columns is bit-identical to the matmul."""
    col = np.sum(A, 0)
    inv = np.zeros_like(col)
    for i in range(A.shape[0]):
        if col[i] > 0:
            inv[i] = col[i] ** (-1)
    return A * inv[None, :]


def normalize_undigraph(A):
    """D^-1/2 A D^-1/2 -- graph.py:487-495 (unused by any strategy, kept for API parity)."""
    col = np.sum(A, 0)
    n = A.shape[0]
    Dn = np.zeros((n, n))
    for i in range(n):
        if col[i] > 0:
            Dn[i, i] = col[i] ** (-0.5)
    return np.dot(np.dot(Dn, A), Dn)


def get_norm(max_hop, hop_dis, num_node, dilation):
    """Column-normalised indicator of {distance 0} U {distance == max_hop} -- graph.py:498-505."""
    adjacency = ((hop_dis == 0) | (hop_dis == max_hop)).astype(np.float64)
    return normalize_digraph(adjacency)


def add_one_distance(adjacency_matrix, A, normalize_adjacency, num_node, kernel_size):
    """Push every non-zero of partitions 1..kernel_size-1 one hop outwards (graph.py:508-518).

    The scan mutates the stack it is reading: column by column (i), row by row (j), a non-zero
    entry [j, i] is overwritten with normalize_adjacency[j, i] and every skeleton neighbour k
    of j (k != i) whose *partition-1* entry [k, i] is still zero receives
    normalize_adjacency[k, i] in the partition being scanned.  Later rows of the same column
    see those writes, so the order is part of the definition."""
    out = np.array(A, dtype=np.float64, copy=True)
    neighbours = [np.flatnonzero(adjacency_matrix[j] == 1) for j in range(num_node)]
    for part in range(1, kernel_size):
        plane = out[part]
        gate = out[1]
        for i in range(num_node):
            for j in range(num_node):
                if plane[j, i] == 0:
                    continue
                plane[j, i] = normalize_adjacency[j, i]
                for k in neighbours[j]:
                    if k != i and gate[k, i] == 0:
                        plane[k, i] = normalize_adjacency[k, i]
    return out


def get_A(A, normalize_adjacency, adjacency_matrix, num_node, kernel_size):
    """graph.py:521-525."""
    return add_one_distance(adjacency_matrix, A, normalize_adjacency, num_node, kernel_size)


def every_symmetric(A, normalize_adjacency, num_node, spatial_symmetric):
    """Append the (one-directional) mirror-joint partition -- graph.py:528-536."""
    mirror = np.zeros((num_node, num_node))
    for i, j in spatial_symmetric:
        mirror[i, j] = normalize_adjacency[i, j]
    return np.concatenate([A, mirror[None]], axis=0)
