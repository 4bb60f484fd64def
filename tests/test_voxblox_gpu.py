"""GPU: voxblox mesh messages (VoxbloxMsgInterface).

Bar: the decoded vertices / colours are bit-identical to the oracle's restatement of
getActiveVertex (voxblox_msg_interface.cpp:43-62), and integrating the packed message gives
exactly what integrating the decoded floats gives — on the GPU and in the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGE = np.float32(0.8)


def voxblox_stream(n_messages, tris_per_block=120, seed=5):
    """A sliding window of voxblox-like blocks (edge 0.8 m): per block u16 coordinates in
    [0, 32767] (one block edge = 65535 / 2 units), triangles up to ~15 cm, colours random."""
    rng = np.random.default_rng(seed)
    for m in range(n_messages):
        blocks = [(bx, by, 0) for bx in range(m, m + 4) for by in range(-1, 2)]
        if m % 3 == 2:
            blocks.append((m + 1, 5, -2))  # an isolated block with negative index
        idx, offsets, xs, ys, zs, cols = [], [0], [], [], [], []
        for b in blocks:
            # geometry is a function of the block so that re-observed blocks repeat it
            brng = np.random.default_rng([seed, b[0] & 0xFFFF, b[1] & 0xFFFF, b[2] & 0xFFFF])
            anchor = brng.integers(0, 32000, size=(tris_per_block, 1, 3))
            tri = np.clip(anchor + brng.integers(0, 6000, size=(tris_per_block, 3, 3)), 0, 32767)
            pts = tri.reshape(-1, 3).astype(np.uint16)
            if len(idx) == 1:
                pts[:3] = [[0, 0, 0], [65535, 65535, 65535], [32767, 1, 65534]]  # range ends
            idx.append(b)
            xs.append(pts[:, 0]); ys.append(pts[:, 1]); zs.append(pts[:, 2])
            cols.append(rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8))
            offsets.append(offsets[-1] + pts.shape[0])
        c = np.concatenate(cols)
        yield (np.array(idx, np.int32), np.array(offsets, np.uint64), np.concatenate(xs),
               np.concatenate(ys), np.concatenate(zs), c[:, 0].copy(), c[:, 1].copy(), c[:, 2].copy(),
               10.0 + m)


def test_decode_matches_oracle(capi, oracle):
    for msg in voxblox_stream(3):
        bi, off, x, y, z, r, g, b, _ = msg
        for edge in (EDGE, np.float32(0.30000001), np.float32(2.5)):
            gx, gc = capi.voxblox_decode(bi, off, x, y, z, r, g, b, edge)
            ox, oc = oracle.voxblox_decode(bi, off, x, y, z, r, g, b, edge)
            assert np.array_equal(gx.view(np.uint32), ox.view(np.uint32))
            assert np.array_equal(gc, oc)
    # alpha is always 255; colour bytes land in PCL's b, g, r, a order
    assert (gc >> 24 == 255).all() and np.array_equal((gc >> 16) & 255, r)
    # empty message
    e = capi.voxblox_decode(np.zeros((0, 3), np.int32), np.zeros(1, np.uint64), *[np.zeros(0, np.uint16)] * 3,
                            *[np.zeros(0, np.uint8)] * 3, EDGE)
    assert e[0].shape == (0, 3)


@pytest.mark.parametrize("kind", ["vcc", "oct", "block"])
def test_packed_integrate_equals_float_integrate(capi, oracle, kind):
    res = 0.05
    if kind == "vcc":
        packed, flat, ora = capi.VccHandle(res, 0), capi.VccHandle(res, 0), oracle.Compression("voxel_clearing", res)
    else:
        packed = capi.OctHandle(res, 0, block=(kind == "block"))
        flat = capi.OctHandle(res, 0, block=(kind == "block"))
        ora = oracle.Compression("block" if kind == "block" else "octree", res)
    for m, msg in enumerate(voxblox_stream(8)):
        bi, off, x, y, z, r, g, b, t = msg
        xyz, rgba = oracle.voxblox_decode(bi, off, x, y, z, r, g, b, EDGE)
        for h in (packed, flat):
            h.prune(t - 3.0)
        ora.prune(t - 3.0)
        got = packed.integrate_voxblox(bi, off, x, y, z, r, g, b, EDGE, t)
        want = flat.integrate(bi, off, xyz, rgba, t)
        if kind == "vcc":
            want = want[0]
        ref = ora.integrate(bi, off, xyz, rgba, t)
        if kind == "vcc":
            assert np.array_equal(got, want)
            assert np.array_equal(got.astype(np.int64), ref[0])
        else:
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
            assert np.array_equal(got[1], ref[1]), f"message {m}: remap kind"
            assert np.array_equal(got[0], ref[0]), f"message {m}: remap index"
    a, b2, o = packed.snapshot(), flat.snapshot(), ora.snapshot()
    for key in ("vertices", "rgba", "stamps", "polygons"):
        assert np.array_equal(a[key], b2[key]), key
        assert np.array_equal(a[key], o[key]), key
    assert a["vertices"].shape[0] > 1000
    for h in (packed, flat):
        h.close()
