"""CPU-side checks of the drop-in boundary: the shared library builds/loads, exports every
symbol include/istgcn_b200.h declares, and the product refuses to run without a GPU."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'istgcn_b200.h')


def _declared():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(istgcn_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib_path():
    import __graft_entry__ as ge
    return ge.build()


def test_header_symbols_are_exported(lib_path):
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r' T (istgcn_[a-z0-9_]+)', out))
    declared = _declared()
    assert len(declared) >= 19
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    from istgcn import _lib
    assert sorted(_lib.SYMBOLS) == declared


def test_library_loads_and_reports_version(lib_path):
    from istgcn import _lib
    lib = _lib.load()
    assert lib.istgcn_version() == 100
    assert isinstance(lib.istgcn_last_error(), bytes)


def test_sm100a_sass_only(lib_path):
    out = subprocess.run(['cuobjdump', '--list-elf', lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_(\d+a?)', out.stdout))
    assert archs == {'100a'}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib_path):
    import net.ist_gcn
    from istgcn import _lib
    model = net.ist_gcn.Model(3, 60, dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), True)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(torch.zeros(1, 3, 8, 25, 2))
    with pytest.raises(RuntimeError, match='CUDA'):
        _lib.call('pool_fwd', torch.zeros(4), torch.zeros(4), 1, 1, 1, 4)


def test_state_dict_layouts_match_oracle():
    """Key names, order and shapes of the drop-in models == the reference layouts that the
    oracle state (strict-loaded into the reference by make_golden.py) encodes."""
    import importlib
    import net.ist_gcn
    import net.st_gcn
    import net.st_gcn_msgcn
    import net.st_gcn_mstcn
    import net.st_gcn_mstcn_1x1
    import net.st_gcnold
    from oracle import model_ref
    assert net.st_gcnold.Model is net.st_gcn.Model
    assert importlib.import_module('net.st_gcn_tanh').Model is net.st_gcn.Model
    cases = [(net.st_gcn.Model, 'st_gcn', dict(layout='ntu-rgb+d', strategy='spatial'), 60),
             (net.st_gcn.Model, 'st_gcn', dict(layout='openpose', strategy='uniform'), 400),
             (net.st_gcn_msgcn.Model, 'st_gcn_msgcn',
              dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60),
             (net.st_gcn_mstcn.Model, 'st_gcn_mstcn', dict(layout='ntu-rgb+d', strategy='spatial'), 60),
             (importlib.import_module('net.st_gcn_mstcn_1x1_deep').Model, 'st_gcn_mstcn_1x1_deep',
              dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60),
             (importlib.import_module('net.st_gcn_deep_msgcn').Model, 'st_gcn_deep_msgcn',
              dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60),
             (importlib.import_module('net.st_gcn_msgcn_new').Model, 'st_gcn_msgcn_new',
              dict(layout='openpose_sym', strategy='spatial_3_sym'), 400),
             (importlib.import_module('net.st_gcn_multi3').Model, 'st_gcn_multi3',
              dict(layout='ntu-rgb+d', strategy='spatial'), 60),
             (importlib.import_module('net.st_gcn_multi3_fix').Model, 'st_gcn_multi3_fix',
              dict(layout='openpose', strategy='spatial'), 400),
             (importlib.import_module('net.st_gcn_only3').Model, 'st_gcn_only3',
              dict(layout='ntu-rgb+d', strategy='spatial'), 60),
             (importlib.import_module('net.st_gcn_learnA').Model, 'st_gcn_learnA',
              dict(layout='ntu-rgb+d', strategy='spatial'), 60),
             (importlib.import_module('net.st_gcn_multi3_fix_3A').Model, 'st_gcn_multi3_fix_3A',
              dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60),
             (importlib.import_module('net.st_gcn_multi3_fix_3A_mstcn').Model, 'st_gcn_multi3_fix_3A_mstcn',
              dict(layout='openpose', strategy='spatial'), 60),
             (net.ist_gcn.Model, 'ist_gcn', dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), 60),
             (net.ist_gcn.Model, 'ist_gcn', dict(layout='openpose_sym', strategy='spatial_3_sym'), 400),
             (net.st_gcn_mstcn_1x1.Model, 'st_gcn_mstcn_1x1',
              dict(layout='ntu-rgb+d_sym', strategy='spatial_sym'), 60)]
    for cls, arch, g_args, ncls in cases:
        m = cls(3, ncls, g_args, True, dropout=0.5)
        st = model_ref.make_state(arch, 3, ncls, m.graph.A, getattr(m.graph, 'A2', None),
                                  getattr(m.graph, 'A3', None))
        sd = m.state_dict()
        assert list(sd.keys()) == list(st.keys())
        for k in st:
            assert sd[k].shape == st[k].shape and sd[k].dtype == st[k].dtype, k
        m.load_state_dict(st, strict=True)
        m2 = cls(3, ncls, g_args, False)
        assert not any(k.startswith('edge_importance') for k in m2.state_dict())
    import net.st_gcn_twostream
    two = net.st_gcn_twostream.Model(3, 60, dict(layout='ntu-rgb+d', strategy='spatial'), True)
    one = list(net.st_gcn.Model(3, 60, dict(layout='ntu-rgb+d', strategy='spatial'), True).state_dict())
    assert list(two.state_dict()) == ['origin_stream.' + k for k in one] + ['motion_stream.' + k for k in one]
    n_params = sum(p.numel() for p in net.ist_gcn.Model(
        3, 60, dict(layout='ntu-rgb+d_sym', strategy='spatial_3_sym'), True).parameters())
    assert n_params == 1100789


def test_sparse_pattern_lists():
    from istgcn.sparse import SparsePattern
    from net.utils.graph import Graph
    import numpy as np
    g = Graph('ntu-rgb+d_sym', 'spatial_3_sym')
    mask = (g.A != 0) | (g.A2 != 0) | (g.A3 != 0)
    p = SparsePattern(mask, 'cpu')
    assert p.nnz == 197 and p.K == 4 and p.V == 25
    k, v, w = np.nonzero(mask)
    flat = p.flat_idx.numpy()
    assert (flat == k * 625 + v * 25 + w).all()
    dptr, dsrc, did = p.dst_ptr.numpy(), p.dst_src.numpy(), p.dst_id.numpy()
    for kw in range(100):
        for j in range(dptr[kw], dptr[kw + 1]):
            assert k[did[j]] * 25 + w[did[j]] == kw and v[did[j]] == dsrc[j]
    sptr, skw, sid = p.src_ptr.numpy(), p.src_kw.numpy(), p.src_id.numpy()
    for vv in range(25):
        for j in range(sptr[vv], sptr[vv + 1]):
            assert v[sid[j]] == vv and k[sid[j]] * 25 + w[sid[j]] == skw[j]
    assert sorted(did) == list(range(197)) and sorted(sid) == list(range(197))
    idn = SparsePattern.identity(18, 'cpu')
    assert idn.nnz == 18 and (idn.dst_src.numpy() == np.arange(18)).all()
