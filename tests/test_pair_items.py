"""Host-side planner of istgcn_gcn_pair_grads (istgcn/sparse.py: SparsePattern.pair_items): the blocks
of the joint-pair pattern it hands to the kernel must own every pattern pair x output column exactly
once and respect the kernel's tensor-memory / shared-memory limits (include/istgcn_b200.h)."""
import numpy as np
import pytest
import torch

from istgcn.sparse import SparsePattern
from net.utils.graph import Graph


@pytest.mark.parametrize('blocks', [None, '0', '1'])
@pytest.mark.parametrize('layout,strategy', [('ntu-rgb+d_sym', 'spatial_3_sym'), ('ntu-rgb+d', 'spatial'),
                                             ('openpose', 'spatial')])
def test_pair_items_cover_pattern_exactly(monkeypatch, layout, strategy, blocks):
    if blocks is None:
        monkeypatch.delenv('ISTGCN_PAIR_BLOCKS', raising=False)
    else:
        monkeypatch.setenv('ISTGCN_PAIR_BLOCKS', blocks)
    g = Graph(layout, strategy)
    A = sum(torch.tensor(getattr(g, n), dtype=torch.float64) for n in ('A', 'A2', 'A3') if hasattr(g, n))
    pat = SparsePattern((A != 0).numpy(), torch.device('cpu'))
    V = pat.V
    in_pattern = np.zeros((V, V), dtype=bool)
    in_pattern[pat._pair_v_host, pat._pair_w] = True
    assert (pat.pair_of.numpy().reshape(V, V) >= 0).tolist() == in_pattern.tolist()
    for cin, cout in ((64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (32, 96)):
        items, ctas, joints = pat.pair_items(cin, cout)
        it, J = items.numpy(), joints.numpy()
        own = np.zeros((V, V, cout), dtype=int)
        for d0, nd, s0, ns, col0, ncw, mask, _ in it:
            assert ncw % 32 == 0 and nd * ncw <= 256 and ns * nd <= 32
            assert ((ns * cin + 127) // 128) * nd * ncw <= 512          # accumulators fit tensor memory
            for vi in range(ns):
                for jd in range(nd):
                    if (int(mask) >> (vi * nd + jd)) & 1:
                        own[J[s0 + vi], J[d0 + jd], col0:col0 + ncw] += 1
        assert (own[in_pattern] == 1).all() and (own[~in_pattern] == 0).all()
        c = ctas.numpy()
        assert set(c[:, 0]) == set(range(len(it)))                      # every item has thread blocks
        for i in range(len(it)):                                        # K-tiles j, j + s, ... partition
            mine = c[c[:, 0] == i]
            assert sorted(mine[:, 1]) == list(range(len(mine))) and (mine[:, 2] == len(mine)).all()
