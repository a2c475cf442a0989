"""CPU oracle for the GNN-propagation -> gather -> MLP -> top-k hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker or the timed CPU
baseline.  Nothing under `deep_cbrs_amar_renaissance_b200/` imports it: the
product path fails loudly if the CUDA library is missing.

What it restates (numpy / scipy, float32, documented summation order):

* graph build      /root/reference/src/data/loaders.py:43-82,
                   src/data/preprocess.py:44-170, src/utilities/math.py:6-56
* layer loop       src/models/gnn.py:74-84, src/layers/reduction.py:15-55,
                   src/layers/lightgcn_conv.py:51-58
* gather + MLP     src/models/basic.py:31-75, src/models/dense.py:4-17,
                   src/models/hybrid.py:72-140, src/layers/fusion.py:49-53
* batching         src/data/datasets.py:190-213,357-366
* top-k            src/utilities/metrics.py:21-34

The arithmetic of GCNConv / GraphSageConv / GATConv / gcn_filter lives in the
third-party package `spektral` (unpinned in requirements.txt:8; era 1.0.x-1.2.0)
on `tensorflow` (unpinned).  Neither is vendored under /root/reference nor
installable here, so those pieces are restated from Spektral's published
algorithm.

PARITY PINNING STATUS
  * graph build / id compaction / batching / gather indices: PINNED - checked
    bit-for-bit against the reference's own numpy/scipy code run unmodified
    under oracle/tf_stub (fixtures in tests/golden/, generator
    tests/golden/make_golden.py).
  * reductions (reduction.py), attention fusion (fusion.py), the DGCF operator recipe and layer
    (dgcf_conv.py), pair-list top-k (metrics.py): PINNED - the reference's own modules run
    unmodified over numpy-backed tf/keras/spektral stand-ins (oracle/tf_np_stub, generator
    tests/golden/make_golden_layers.py, vectors tests/golden/layers/).
  * model wiring end to end (SequentialGNN loop, family builders, lookups, BasicRS, HybridCBRS in all
    modes, dense builders): PINNED - src/models/*.py run unmodified over the same stand-ins
    (tests/golden/make_golden_models.py, tests/golden/models/); leaves are stand-ins.
  * Spektral layer arithmetic (GCNConv, GraphSageConv, GATConv, gcn_filter), Keras Dense, the
    training step: PARITY UNPINNED - the
    reference ships no tests, golden vectors or fixtures and TF/Spektral cannot
    run here.  Pinned only by known-answer parameter counts (doc.pdf Table 17)
    and algebraic identities (tests/test_oracle_identities.py).
"""
