"""Oracle of the mesh -> graph index bookkeeping (oracle/frontend_oracle.h): hand-computed cases
of MeshFrontendInterface::updateMeshToGraphMappings (mesh_frontend_interface.cpp:197-211) and
consistency with the remapping the oracle compressors maintain themselves."""
import numpy as np

from kimera_pgmo_b200 import synth
from oracle import ffi

A, B = (0, 0, 0), (1, 0, 0)


def pairs(o):
    m, g = o.mesh_to_graph()
    return dict(zip(m.tolist(), g.tolist()))


def test_first_insert_wins_and_stale_entries_stay():
    o = ffi.MeshToGraph()
    # block A, 3 msg vertices: two of them on mesh vertex 10 with different graph vertices
    o.update([A], [0, 3], [10, 10, 11], None, [5, 6, 6], None)
    assert pairs(o) == {10: 5, 11: 6}  # msg index 0 is walked before 1
    # a smaller version of A: entry 2 (mesh 11 -> graph 6) is never erased
    o.update([A], [0, 2], [12, 10], None, [7, 8], None)
    assert pairs(o) == {10: 5, 11: 6, 12: 7}  # insert() keeps 10 -> 5
    assert o.block_mapping(A, 0).tolist() == [12, 10, 11]
    assert o.block_mapping(A, 1).tolist() == [7, 8, 6]
    assert o.block_mapping(B, 0) is None


def test_write_kinds_and_missing_graph_entries():
    o = ffi.MeshToGraph()
    # kind 0 = no write: msg vertex 1 has no graph entry, so mesh vertex 21 gets no pair (:203)
    o.update([A, B], [0, 2, 4], [20, 21, 22, 23], [1, 1, 1, 1], [1, 0, 2, 3], [1, 0, 2, 1])
    assert pairs(o) == {20: 1, 22: 2, 23: 3}
    assert o.block_mapping(A, 1).tolist() == [1, -1]
    # kind 2 = insert-if-absent: A[0] keeps graph 1, A[1] receives 9; a stale mesh entry whose
    # msg index receives its graph entry only now is paired now (21 -> 9)
    o.update([A], [0, 2], [0, 0], [0, 0], [8, 9], [2, 2])
    assert o.block_mapping(A, 1).tolist() == [1, 9]
    assert pairs(o) == {20: 1, 21: 9, 22: 2, 23: 3}


def test_block_order_of_the_message_decides_between_blocks():
    o1, o2 = ffi.MeshToGraph(), ffi.MeshToGraph()
    # the same mesh vertex 30 reached from both blocks with different graph vertices
    o1.update([A, B], [0, 1, 2], [30, 30], None, [1, 2], None)
    o2.update([B, A], [0, 1, 2], [30, 30], None, [2, 1], None)
    assert pairs(o1) == {30: 1}
    assert pairs(o2) == {30: 2}


def test_block_maps_equal_the_compressors_own_remapping():
    """The (index, kind) records of integrate(), replayed into dense per-block arrays, reproduce
    the HashedIndexMapping the oracle compressors keep (the reference's msg_to_*_idx_)."""
    full, graph, m2g = ffi.Compression("voxel_clearing", 0.05), ffi.Compression("octree", 0.4), ffi.MeshToGraph()
    seen = {}
    for blocks, offsets, xyz, rgba, t in synth.compression_stream(6, 3000, n_blocks=5, seed=9):
        full.prune(t - 3.0)
        graph.prune(t - 3.0)
        mi, mk, _ = full.integrate(blocks, offsets, xyz, rgba, t)
        gi, gk, _ = graph.integrate(blocks, offsets, xyz, rgba, t)
        m2g.update(blocks, offsets, mi, mk, gi, gk)
        for b, blk in enumerate(np.asarray(blocks).reshape(-1, 3)):
            seen[tuple(int(v) for v in blk)] = max(seen.get(tuple(int(v) for v in blk), 0),
                                                   int(offsets[b + 1] - offsets[b]))
    assert len(seen) > 5
    n_pairs = 0
    for blk, length in seen.items():
        for which, comp in ((0, full), (1, graph)):
            dense = m2g.block_mapping(blk, which)
            assert len(dense) == length
            assert int((dense >= 0).sum()) == comp.remap_block_size(blk)
            for i in range(0, length, 7):
                assert dense[i] == comp.remap_lookup(blk, i)
    m, g = m2g.mesh_to_graph()
    n_pairs = len(m)
    assert n_pairs > 100
    assert np.all(np.diff(m.astype(np.int64)) > 0)
    assert int(g.max()) < graph.sizes()[0] and int(m.max()) < full.sizes()[0]
