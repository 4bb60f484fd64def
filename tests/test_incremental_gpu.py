"""GPU parity of the incremental deformation path: DeformationGraph::deformPoints<CloudIn,CloudOut>
(kimera_pgmo/include/kimera_pgmo/deformation_graph.h:808-870) = findStartIndex (:715-746) +
fillPreviousPoints (:748-760) + predeformPoints (:779-806) + deformation::deformPoints over the
remaining indices + cacheNewPoints (:762-777).

The device work goes through the C ABI (pgmo_deform_predeform, pgmo_deform_points with a subset);
the few lines of host bookkeeping below restate what the C++ shim does
(include/kimera_pgmo_b200/deformation_graph.h) so that the sequence can be driven from Python.
Expected values: the oracle's DeformationGraphMeshPath (oracle/deform_oracle.cpp) on the same
inputs.  Bar: bit-identical floats, identical start indices and associations.
"""
import numpy as np
import pytest

from kimera_pgmo_b200 import synth

pytestmark = pytest.mark.gpu

U64 = np.uint64


def _upper_bound(stamps, value):
    """std::upper_bound(gen.begin(), gen.end(), value, [](v, i) { return v < stamp[i]; })"""
    first, count = 0, len(stamps)
    while count > 0:
        step = count // 2
        it = first + step
        if not (value < stamps[it]):
            first = it + 1
            count -= step + 1
        else:
            count = step
    return first


class GpuGraphPath:
    """Host bookkeeping of deformPoints<CloudIn,CloudOut> around the C ABI."""

    def __init__(self, capi):
        self.capi = capi
        self.h = capi.DeformHandle(0)
        self.ctrl_stamps = None
        self.cache = None           # last_calculated_vertices_[prefix]
        self.recalculate = False

    def set_controls(self, ctrl_xyz, ctrl_stamps, valid, poses):
        self.h.set_controls(ctrl_xyz, ctrl_stamps, valid, poses)
        self.ctrl_stamps = np.asarray(ctrl_stamps, dtype=U64)

    def find_start(self, stamps, tol_t, hint):
        if self.cache is None:                      # :719-723
            return 0
        if self.recalculate:                        # :725-728
            return 0
        if hint >= 0:                               # :730-732
            return hint
        # :734-736 unsigned: back - stampFromSec(tol) wraps, max(0, .) never clips
        min_stamp = U64((int(self.ctrl_stamps[-1]) - int(self.capi.stamp_from_sec(tol_t))) % (1 << 64))
        return min(_upper_bound(stamps, min_stamp), self.cache[0].shape[0])  # :738-745

    def deform(self, x, y, z, stamps, k, tol_t, graph_indices=None, hint=-1, stride=8):
        n = x.shape[0]
        start = self.find_start(stamps, tol_t, hint)
        out = [x.copy(), y.copy(), z.copy()]        # `vertices` starts as a copy of the input
        for o, c in zip(out, self.cache or ()):     # fillPreviousPoints
            o[:start] = c[:start]
        assoc = np.full((n, stride), -1, dtype=np.int64)
        cnt = np.zeros(n, dtype=np.int64)
        subset = None
        if start != 0:
            if graph_indices is not None:
                subset = self.h.predeform(x, y, z, graph_indices, start, out)
            else:
                subset = np.arange(start, n, dtype=U64)
        r = self.h.deform_points(x, y, z, stamps, k=k, tol_t=tol_t, subset=subset, out=out,
                                 want_assoc=True, assoc_stride=stride)
        if r["rc"] == 0:
            a = r["assoc"].astype(np.int64)
            a[a == 0xFFFFFFFF] = -1
            done = r["processed"].astype(bool)
            rows = np.arange(n) if subset is None else subset.astype(np.int64)
            if subset is None:
                # start == 0: entry e of the map = e-th processed vertex (reference :858-859)
                e = np.flatnonzero(done)
                assoc[: e.shape[0]] = a[e]
                cnt[: e.shape[0]] = r["assoc_cnt"][e]
            else:
                # the reference scatters map entry i to indices[i] (:861-865); entries exist only
                # for processed positions, in order
                e = np.flatnonzero(done)
                assoc[rows[: e.shape[0]]] = a[e]
                cnt[rows[: e.shape[0]]] = r["assoc_cnt"][e]
        if self.cache is None or self.cache[0].shape[0] != n:   # cacheNewPoints: resize keeps the prefix
            old = self.cache
            self.cache = [np.zeros(n, dtype=np.float32) for _ in range(3)]
            if old is not None:
                m = min(n, old[0].shape[0])
                for c, o in zip(self.cache, old):
                    c[:m] = o[:m]
        for c, o in zip(self.cache, out):
            c[start:] = o[start:]
        self.recalculate = False
        return dict(start=start, x=out[0], y=out[1], z=out[2], assoc=assoc, assoc_cnt=cnt)


def _compare(gpu, ref, check_assoc=True):
    assert gpu["start"] == ref["start"]
    for ax in "xyz":
        np.testing.assert_array_equal(gpu[ax], ref[ax])
    if check_assoc:
        np.testing.assert_array_equal(gpu["assoc_cnt"], ref["assoc_cnt"])
        np.testing.assert_array_equal(gpu["assoc"], ref["assoc"].astype(np.int64))


def _session(capi, oracle, case, splits, tol_t, k=4, graph_indices=None, hints=None, recalc_at=(),
             t_check_assoc=True):
    """Replays a growing session: step s sees the first splits[s] = (n_vertices, n_controls)."""
    gpu, ora = GpuGraphPath(capi), oracle.GraphPath()
    for s, (n_v, n_c) in enumerate(splits):
        ctrl = (case["ctrl_xyz"][:n_c], case["ctrl_stamps"][:n_c], case["valid"][:n_c],
                case["poses_per_step"][s][:n_c])
        gpu.set_controls(*ctrl)
        ora.set_controls("a", *ctrl)
        if s in recalc_at:
            gpu.recalculate = True
            ora.set_recalculate(True)
        gi = None if graph_indices is None else graph_indices[:n_v]
        hint = -1 if hints is None else hints[s]
        args = (case["x"][:n_v], case["y"][:n_v], case["z"][:n_v], case["stamps"][:n_v], k, tol_t)
        g = gpu.deform(*args, graph_indices=gi, hint=hint)
        r = ora.deform("a", *args, graph_indices=gi, hint=hint)
        _compare(g, r, t_check_assoc)
        yield s, g


def _case(n_v, n_c, seed, t0_ns=synth.T0_NS, steps=4):
    case = synth.deformation_case(n_v, n_c, seed=seed, t0_ns=t0_ns, invalid_fraction=0.02)
    rng = np.random.default_rng(seed)
    case["poses_per_step"] = []
    for s in range(steps):
        p = case["poses"].copy()
        p[:, 9:] += rng.normal(0, 0.02 * (s + 1), p[:, 9:].shape)
        case["poses_per_step"].append(p)
    return case


def test_growing_mesh_without_graph_indices(capi, oracle):
    case = _case(60_000, 1_200, seed=41)
    splits = [(30_000, 600), (45_000, 900), (45_000, 900), (60_000, 1_200)]
    starts = [g["start"] for _, g in _session(capi, oracle, case, splits, 10.0)]
    assert starts[0] == 0 and 0 < starts[1] <= 30_000 and 0 < starts[3] <= 45_000


def test_graph_indices_mix_and_cache_reuse(capi, oracle):
    case = _case(60_000, 1_200, seed=42)
    rng = np.random.default_rng(3)
    n_v, n_c = 60_000, 1_200
    gi = rng.integers(0, n_c, n_v).astype(np.int32)
    gi[rng.random(n_v) < 0.3] = -1                      # no graph vertex yet: falls back to knn
    gi[rng.random(n_v) < 0.05] = n_c + 50               # key not in the values: falls back too
    splits = [(20_000, 500), (40_000, 900), (60_000, 1_200), (60_000, 1_200)]
    got = list(_session(capi, oracle, case, splits, 10.0, graph_indices=gi))
    assert got[1][1]["start"] > 0 and got[3][1]["start"] > 0


def test_start_index_hint_and_recalculate(capi, oracle):
    case = _case(40_000, 800, seed=43)
    splits = [(20_000, 400), (30_000, 600), (40_000, 800), (40_000, 800)]
    rng = np.random.default_rng(5)
    gi = rng.integers(-1, 800, 40_000).astype(np.int32)
    got = list(_session(capi, oracle, case, splits, 10.0, graph_indices=gi,
                        hints=[-1, 12_345, 25_000, -1], recalc_at=(3,)))
    assert [g["start"] for _, g in got] == [0, 12_345, 25_000, 0]


def test_unsigned_wrap_of_the_start_stamp(capi, oracle):
    # stamps from 0 with tol_t larger than the newest control stamp: back - tol wraps
    # (deformation_graph.h:734-736), upper_bound runs off the end, start = cache size
    case = _case(30_000, 600, seed=44, t0_ns=0)
    splits = [(15_000, 300), (30_000, 600), (30_000, 600)]
    got = list(_session(capi, oracle, case, splits, 1.0e3))
    assert got[1][1]["start"] == 15_000 and got[2][1]["start"] == 30_000


@pytest.mark.parametrize("k", [2, 6])
def test_other_k(capi, oracle, k):
    case = _case(24_000, 480, seed=45)
    splits = [(12_000, 240), (24_000, 480)]
    list(_session(capi, oracle, case, splits, 10.0, k=k))
