"""CPU oracle for the IST-GCN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and only as the checker
(or, for the CPU baseline, as the thing timed on the host cores).  Nothing under
``ist-gcn_b200/`` imports it; the product path raises if its CUDA library is missing.

Contents
  graph_ref.py   loop-for-loop NumPy restatement of the reference graph construction
                 (reference: net/utils/graph.py).
  model_ref.py   functional plain-PyTorch fp32/fp64 restatement of the reference networks
                 (reference: net/utils/tgcn.py, net/utils/inceptionv2_gcn.py,
                 net/st_gcnold.py, net/st_gcn_msgcn.py, net/st_gcn_mstcn.py,
                 net/st_gcn_mstcn_1x1.py, net/st_gcn_twostream.py,
                 processor/recognition.py:31-44,152-176).
  refload.py     imports the real reference from /root/reference next to the repo's own
                 ``net`` package (only possible in the build container; used by
                 tests/golden/make_golden.py and by the live-parity CPU tests).

Pinning.  The reference ships no tests, fixtures or known-answer vectors for this path
(SURVEY.md section 4, section 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
tests/golden/*.npz were generated in the build container by tests/golden/make_golden.py,
which imports the reference modules from /root/reference, runs them on seeded inputs and
stores inputs, weights, logits and gradients.  tests/test_oracle.py checks graph_ref and
model_ref against those fixtures (and against the live reference when it is mounted), and
tests/test_graph.py checks the SHA-256 table of SURVEY.md App. A.  Third-party arithmetic
underneath the reference (torch conv/einsum/batch_norm, unpinned in requirements.txt) is
pinned only to the behaviour of the installed torch 2.11 / numpy 2.3.
"""
