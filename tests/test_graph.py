"""The product's skeleton tables (kinetic-gan_b200/models/init_gan: graph_ntu / Graph_h36m over skeleton.SkeletonGraph, frozen
coarsening + vectorised derivation) against the tables of the UNMODIFIED reference classes (models/init_gan/graph_ntu.py,
graph_h36m.py - networkx-driven), dumped by tests/golden/make_golden.py into tests/golden/graph_*.npz: adjacency partitions per
level (bit-equal), kept-joint maps, edges, centres, and the up-sampling neighbourhoods."""
import numpy as np
import pytest

import kgan_b200 as kgan
from helpers import load_golden
from oracle.graph import SkeletonTables


@pytest.mark.parametrize("name,cls", [("ntu", "graph_ntu"), ("h36m", "Graph_h36m")])
def test_product_graph_tables_equal_reference(name, cls):
    gold = load_golden("graph_" + name)
    g = getattr(kgan, cls)() if hasattr(kgan, cls) else None
    if g is None:
        from importlib import import_module
        mod = import_module("kinetic-gan_b200.models.init_gan." + ("graph_ntu" if name == "ntu" else "graph_h36m"))
        g = getattr(mod, cls)()
    assert list(g.num_node) == list(gold["num_node"]) and list(g.center) == list(gold["center"])
    assert g.lvls == len(gold["num_node"])
    for lvl in range(g.lvls):
        A = np.asarray(g.As[lvl])
        assert A.dtype == np.float64 and A.shape == gold["As%d" % lvl].shape
        assert np.array_equal(A, gold["As%d" % lvl]), lvl                         # bit-equal adjacency partitions
        assert np.array_equal(np.asarray(g.map[lvl]), gold["map%d" % lvl]), lvl
    for lvl in range(g.lvls - 1):
        n = int(gold["mapping%d_len" % lvl])
        assert len(g.mapping[lvl]) == n
        for i in range(n):
            assert [int(v) for v in g.mapping[lvl][i]] == [int(v) for v in gold["mapping%d_%d" % (lvl, i)]], (lvl, i)
    # the oracle's restatement agrees too (it is what the parity tests run against)
    t = SkeletonTables(name)
    for lvl in range(g.lvls):
        assert np.array_equal(np.asarray(t.As[lvl]), gold["As%d" % lvl])
