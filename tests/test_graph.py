"""Graph parity: the product Graph, the oracle restatement, the golden
arrays generated from the reference, the SHA-256 table of SURVEY.md App. A and -- when the
reference tree is mounted -- the live reference must agree BIT FOR BIT."""
import hashlib
import itertools
import os

import numpy as np
import pytest

from net.utils.graph import Graph
from oracle import graph_ref, refload

# first 16 hex digits of sha256(float64 bytes): SURVEY.md App. A (generated from the reference)
SHA = {
    ('ntu-rgb+d', 'uniform'): ('9e160f59c73e105c',),
    ('ntu-rgb+d', 'distance'): ('21077e85b2df1c35',),
    ('ntu-rgb+d', 'spatial'): ('37b570e92c97dcd0',),
    ('ntu-rgb+d', 'spatial_half'): ('37b570e92c97dcd0',),
    ('ntu-rgb+d', 'spatial_3'): ('37b570e92c97dcd0', '6851f204dcc252ee', '94906017d44e8d85'),
    ('ntu-rgb+d', 'spatial_sym'): ('fc4e70208770e1b1',),
    ('ntu-rgb+d', 'spatial_3_sym'): ('fc4e70208770e1b1', 'c8b28a7ceca41b6f', '68dc059a6e8faff3'),
    ('ntu-rgb+d', 'openpose_gravity'): ('509624a0b6059c5a',),
    ('ntu-rgb+d_sym', 'uniform'): ('7e949c0059a34983',),
    ('ntu-rgb+d_sym', 'distance'): ('b4a81c4f25b4ac7c',),
    ('ntu-rgb+d_sym', 'spatial'): ('5ff4942121d77524',),
    ('ntu-rgb+d_sym', 'spatial_3'): ('5ff4942121d77524', '61aaf7267034ecaf', '9733f31a280dc6a4'),
    ('ntu-rgb+d_sym', 'spatial_sym'): ('490d4833977930da',),
    ('ntu-rgb+d_sym', 'spatial_3_sym'): ('21ac2f8712cbbd6c', '1e6b9348b480274d', '88049eed55b4b6d6'),
    ('ntu-rgb+d_sym', 'openpose_gravity'): ('020c7e536a567d91',),
    ('openpose', 'uniform'): ('956d80a55d830d2a',),
    ('openpose', 'distance'): ('f346773912192718',),
    ('openpose', 'spatial'): ('229381604220fcf2',),
    ('openpose_sym', 'spatial_3'): ('229381604220fcf2', '23a6c349d860c68c', '212b6f90ebf2f216'),
    ('openpose_sym', 'spatial_sym'): ('e1e10c4b931aff85',),
    ('openpose_sym', 'spatial_3_sym'): ('5406be2c571d43c6', 'c7b7be7f80e2d01a', 'fff80b37876db0a6'),
    ('ntu-rgb+d_half', 'uniform'): ('b9de3643a08ff368',),
    ('ntu-rgb+d_half', 'distance'): ('4c2bdb6a30cd8e32',),
    ('ntu-rgb+d_half', 'spatial'): ('5df4915d6d3b6b73',),
    ('ntu-rgb+d_half', 'spatial_3'): ('5df4915d6d3b6b73', '83b56c8ae005d43f', 'a10279d387746e07'),
    ('ntu-rgb+d_half', 'spatial_3_sym'): ('32aa1dce2ff89108', '670b3088855f45de', 'ace076f4776e4370'),
}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize('layout,strategy', sorted(SHA))
def test_sha_table(layout, strategy):
    g = Graph(layout=layout, strategy=strategy)
    got = tuple(_sha(getattr(g, n)) for n in ('A', 'A2', 'A3') if hasattr(g, n))
    assert got == SHA[(layout, strategy)]
    assert g.A.dtype == np.float64


def test_golden_arrays_product_and_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, 'graphs.npz'))
    seen = 0
    for key in z.files:
        layout, strategy, name = key.split('|')
        ref = z[key]
        prod = getattr(Graph(layout=layout, strategy=strategy), name)
        orc = graph_ref.build(layout, strategy)[name]
        assert prod.shape == ref.shape and prod.tobytes() == ref.tobytes(), key
        assert orc.shape == ref.shape and orc.tobytes() == ref.tobytes(), key
        seen += 1
    assert seen == 55


def test_facts_from_survey():
    g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
    assert [int((g.A[k] != 0).sum()) for k in range(4)] == [25, 24, 24, 10]
    assert [int((g.A2[k] != 0).sum()) for k in range(4)] == [25, 20, 34, 0]
    assert [int((g.A3[k] != 0).sum()) for k in range(4)] == [25, 16, 44, 0]
    union = (g.A != 0) | (g.A2 != 0) | (g.A3 != 0)
    assert int(union.sum()) == 197
    k = Graph('openpose_sym', 'spatial_3_sym')
    assert int(((k.A != 0) | (k.A2 != 0) | (k.A3 != 0)).sum()) == 152
    s = Graph('ntu-rgb+d_sym', 'spatial_sym')
    nz = {(int(i), int(j)): float(s.A[3, i, j]) for i, j in zip(*np.nonzero(s.A[3]))}
    assert nz == {(8, 4): 0.2, (16, 12): 0.25}
    assert Graph('ntu-rgb+d', 'spatial_sym').A.shape == (4, 25, 25)
    assert not Graph('ntu-rgb+d', 'spatial_sym').A[3].any()


def test_error_behaviour():
    with pytest.raises(ValueError, match='Do Not Exist This Layout'):
        Graph(layout='nope')
    with pytest.raises(ValueError, match='Do Not Exist This Strategy'):
        Graph(layout='openpose', strategy='spatial_gravity')
    for layout in ('openpose_gravity', 'ntu-rgb+d_gravity', 'ntu_edge'):
        with pytest.raises(AttributeError):          # reference never defines spatial_symmetric
            Graph(layout=layout, strategy='spatial')
    with pytest.raises(IndexError):
        Graph(layout='openpose', strategy='openpose_gravity')


@pytest.mark.skipif(not refload.available(), reason='reference tree not mounted')
def test_live_reference_every_combination():
    ref_mod = refload.load('net.utils.graph')
    layouts = ['openpose', 'openpose_sym', 'openpose_gravity', 'ntu-rgb+d', 'ntu-rgb+d_sym',
               'ntu-rgb+d_half', 'ntu-rgb+d_gravity', 'ntu_edge', 'bogus']
    strategies = ['uniform', 'distance', 'spatial', 'spatial_half', 'openpose_gravity',
                  'ntu-rgb+d_gravity', 'spatial_3', 'spatial_sym', 'spatial_3_sym', 'bogus']
    ran = 0
    for layout, strategy in itertools.product(layouts, strategies):
        def attempt(cls):
            try:
                return 'ok', cls(layout=layout, strategy=strategy)
            except Exception as exc:            # noqa: BLE001 - the exception type IS the contract
                return type(exc).__name__, None
        r, p = attempt(ref_mod.Graph), attempt(Graph)
        assert r[0] == p[0], (layout, strategy)
        if r[0] != 'ok':
            continue
        ran += 1
        for name in ('A', 'A2', 'A3', 'hop_dis', 'hop_dis_sym', 'hop_dis23', 'adjacency_matrix'):
            a, b = getattr(r[1], name, None), getattr(p[1], name, None)
            assert (a is None) == (b is None), (layout, strategy, name)
            if a is not None:
                assert a.shape == b.shape and a.tobytes() == b.tobytes(), (layout, strategy, name)
        assert r[1].edge == p[1].edge and r[1].center == p[1].center
    assert ran == 37
