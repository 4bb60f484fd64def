"""pgmo_m2g_* (mesh -> graph index bookkeeping in HBM) against oracle/frontend_oracle.h:
bit-exact index work (mesh_frontend_interface.cpp:100-115, :197-211)."""
import numpy as np
import pytest

from kimera_pgmo_b200 import capi, synth
from oracle import ffi

pytestmark = pytest.mark.gpu


def same_state(gpu, ora, blocks):
    gm, gg = gpu.mesh_to_graph()
    om, og = ora.mesh_to_graph()
    assert np.array_equal(gm, om) and np.array_equal(gg, og)
    for blk in blocks:
        for which in (0, 1):
            assert np.array_equal(gpu.block_mapping(blk, which), ora.block_mapping(blk, which))


def test_hand_case():
    gpu = capi.MeshToGraphHandle(0)
    gpu.update([(0, 0, 0)], [0, 3], [10, 10, 11], None, [5, 6, 6], None)
    gpu.update([(0, 0, 0)], [0, 2], [12, 10], None, [7, 8], None)
    m, g = gpu.mesh_to_graph()
    assert dict(zip(m.tolist(), g.tolist())) == {10: 5, 11: 6, 12: 7}
    assert gpu.block_mapping((0, 0, 0), 0).tolist() == [12, 10, 11]
    assert gpu.block_mapping((0, 0, 0), 1).tolist() == [7, 8, 6]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_records(seed):
    """Blocks that grow and shrink, all three write kinds, heavy index collisions, repeated
    blocks across messages in changing order."""
    rng = np.random.default_rng(seed)
    gpu, ora = capi.MeshToGraphHandle(0), ffi.MeshToGraph()
    universe = [(int(x), int(y), 0) for x in range(-3, 4) for y in range(2)]
    seen = set()
    for msg in range(25):
        nb = int(rng.integers(1, 9))
        pick = rng.permutation(len(universe))[:nb]
        blocks = [universe[i] for i in pick]
        sizes = rng.integers(0, 40, nb) * 3
        if msg % 5 == 4:
            sizes[0] = 3000  # a large version of some block: its segment moves
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        n = int(offsets[-1])
        span = 50 + 400 * msg  # the mesh grows; low indices keep colliding
        mi = rng.integers(0, span, n).astype(np.int64)
        gi = rng.integers(0, 64, n).astype(np.int64)
        mk = rng.choice([0, 1, 2], n, p=[0.1, 0.6, 0.3]).astype(np.uint8) if msg % 3 else None
        gk = rng.choice([0, 1, 2], n, p=[0.3, 0.4, 0.3]).astype(np.uint8)
        gpu.update(blocks, offsets, mi, mk, gi, gk)
        ora.update(blocks, offsets, mi, mk, gi, gk)
        seen.update(blocks)
        same_state(gpu, ora, blocks)
    same_state(gpu, ora, sorted(seen))
    span_dev = gpu.device_view()[1]
    assert span_dev >= 50 + 400 * 24


def test_frontend_pipeline():
    """Both compressors on the GPU feed the GPU bookkeeping; the oracle chain is the reference."""
    g_full, g_graph, g_m2g = capi.VccHandle(0.05, 0), capi.OctHandle(0.4, 0), capi.MeshToGraphHandle(0)
    o_full, o_graph, o_m2g = ffi.Compression("voxel_clearing", 0.05), ffi.Compression("octree", 0.4), ffi.MeshToGraph()
    seen = set()
    for blocks, offsets, xyz, rgba, t in synth.compression_stream(8, 6000, n_blocks=7, seed=4):
        for c in (g_full, g_graph, o_full, o_graph):
            c.prune(t - 3.0)
        remap, _ = g_full.integrate(blocks, offsets, xyz, rgba, t)
        gi, gk = g_graph.integrate(blocks, offsets, xyz, rgba, t)
        g_m2g.update(blocks, offsets, remap.view(np.int64), None, gi, gk)
        omi, omk, _ = o_full.integrate(blocks, offsets, xyz, rgba, t)
        ogi, ogk, _ = o_graph.integrate(blocks, offsets, xyz, rgba, t)
        o_m2g.update(blocks, offsets, omi, omk, ogi, ogk)
        seen.update(tuple(int(v) for v in b) for b in np.asarray(blocks).reshape(-1, 3))
    same_state(g_m2g, o_m2g, sorted(seen))
    m, g = g_m2g.mesh_to_graph()
    assert len(m) > 1000


def test_errors():
    gpu = capi.MeshToGraphHandle(0)
    with pytest.raises(capi.PgmoError):
        gpu.block_mapping((9, 9, 9), 0)  # the reference's .at() throws
    with pytest.raises(capi.PgmoError):
        gpu.update([(0, 0, 0)], [0, 3], [1, -2, 3], None, [0, 0, 0], None)
    with pytest.raises(capi.PgmoError):
        gpu.update([(0, 0, 0)], [0, 3], [1, 2, 3], [1, 3, 1], [0, 0, 0], None)
    gpu.update([], [0], [], None, [], None)
    m, g = gpu.mesh_to_graph()
    assert len(m) == 0
