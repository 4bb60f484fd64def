"""Pins the oracle's restatement of pcl::octree::OctreePointCloudSearch against REAL PCL.

SURVEY.md §8(c) (i)-(iii) — tie order / K-boundary winners of nearestKSearch, leaf count vs point
count and multi-point voxel deletion, bounding-box growth and the greedy approxNearestSearch —
are defined in this repository by oracle/pcl_octree_restated.h, because PCL does not exist in the
build image.  oracle/pcl_fixtures/make_fixtures.cpp runs ~4900 such operations; compiled with
PCL (oracle/pcl_fixtures/CMakeLists.txt) it records PCL's answers, compiled with the
restatement (make -C oracle) it replays a recording and compares result by result.

One command on any machine with PCL turns "parity unpinned" into a checked fixture:

    cmake -S oracle/pcl_fixtures -B /tmp/pclfix && cmake --build /tmp/pclfix && \\
        /tmp/pclfix/make_fixtures generate tests/golden/pcl_fixtures.txt
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "pcl_fixtures", "fixtures_restated")
GOLDEN = os.path.join(ROOT, "tests", "golden", "pcl_fixtures.txt")
HOW = ("no PCL recording committed.  On a machine with PCL: cmake -S oracle/pcl_fixtures -B /tmp/pclfix "
       "&& cmake --build /tmp/pclfix && /tmp/pclfix/make_fixtures generate tests/golden/pcl_fixtures.txt")


@pytest.fixture(scope="module")
def checker():
    if not os.path.exists(BIN):
        p = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "pcl_fixtures/fixtures_restated"],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stdout + p.stderr
    return BIN


def test_recording_round_trips_through_the_restatement(checker, tmp_path):
    """Format, generator and replayer work: a recording made by the restatement replays on the
    restatement without a difference, and covers every case family."""
    dump = tmp_path / "fixtures.txt"
    p = subprocess.run([checker, "generate", str(dump)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    text = dump.read_text()
    cases = [line.split()[1] for line in text.splitlines() if line.startswith("case ")]
    for family in ("ties_res1", "ties_res005", "voxels_shared", "voxels_crowded_leaf", "growth_res1",
                   "growth_res005", "approx_res1", "reanchor_res005", "random_window_res1"):
        assert family in cases
    assert sum(1 for line in text.splitlines() if "|" in line) > 4000
    p = subprocess.run([checker, "check", str(dump)], capture_output=True, text=True)
    assert p.returncode == 0 and " 0 mismatches" in p.stdout, p.stdout + p.stderr
    # a corrupted result is noticed
    lines = text.splitlines()
    i = next(n for n, line in enumerate(lines) if line.startswith("leaves |"))
    lines[i] = lines[i] + "9"
    bad = tmp_path / "bad.txt"
    bad.write_text("\n".join(lines) + "\n")
    p = subprocess.run([checker, "check", str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and "MISMATCH" in p.stderr


def test_restatement_matches_real_pcl(checker):
    if not os.path.exists(GOLDEN):
        pytest.skip(HOW)
    head = open(GOLDEN).read(400)
    assert "# backend pcl" in head, "tests/golden/pcl_fixtures.txt was not recorded with real PCL"
    p = subprocess.run([checker, "check", GOLDEN], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
