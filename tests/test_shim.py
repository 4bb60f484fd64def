"""The header-only C++ shim (include/kimera_pgmo_b200/) mirrors the reference's C++ API.
CPU: it compiles against the stand-in PCL/GTSAM types and links libpgmo_b200.so.
GPU: tests/cpp/test_shim.cpp replays the reference's own unit tests through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "tests", "test_shim")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    lib_dir = os.path.join(ROOT, "kimera_pgmo_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I",
           os.path.join(ROOT, "include"), "-o", BIN,
           os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), "-L", lib_dir, "-lpgmo_b200",
           f"-Wl,-rpath,{lib_dir}"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]


def test_shim_compiles_and_links(capi):
    _build()
    p = subprocess.run([BIN], capture_output=True, text=True)
    # without a GPU the binary must refuse to run (exit 77), never fall back to a CPU path
    assert p.returncode in (0, 77), p.stdout + p.stderr


def test_stdsort_restatement_matches_toolchain():
    exe = os.path.join(ROOT, "build", "tests", "test_stdsort")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "kimera_pgmo_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_stdsort.cpp")], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_shim_host_side_helpers(capi):
    """parallelFor / PinnedBlock arithmetic of the shim: pure host code, no device call."""
    exe = os.path.join(ROOT, "build", "tests", "test_shim_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    lib_dir = os.path.join(ROOT, "kimera_pgmo_b200", "lib")
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread", "-I",
                        os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(ROOT, "tests", "cpp", "test_shim_host.cpp"), "-L", lib_dir,
                        "-lpgmo_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_shim_host_logic_against_oracle_backed_double():
    """Every test of tests/cpp/test_shim.cpp (the reference's unit tests re-typed against the
    shim + the large threaded-staging case) with tests/cpp/mock_{deform,compression}_backend.cpp
    — TEST DOUBLES of the C ABI on top of the CPU oracle — linked instead of libpgmo_b200.so:
    checks the shim's host logic (start index, cache, predeform scatter, staging, `remapping`
    maintenance, MeshDelta reconstruction, block filters, getters) without a GPU.  The product
    library has no CPU path; the GPU suite runs the same file against it."""
    exe = os.path.join(ROOT, "build", "tests", "test_shim_hostlogic")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    oracle_dir = os.path.join(ROOT, "oracle")
    p = subprocess.run(["make", "-C", oracle_dir, "-j4"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    p = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread",
                        "-DSHIM_TEST_HOST_DOUBLE", "-I", os.path.join(ROOT, "include"),
                        "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                        os.path.join(ROOT, "tests", "cpp", "mock_deform_backend.cpp"),
                        os.path.join(ROOT, "tests", "cpp", "mock_compression_backend.cpp"),
                        "-L", oracle_dir, "-loracle", f"-Wl,-rpath,{oracle_dir}"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, "--octree"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 failures" in p.stdout
    # page-locking refused (e.g. the locked-memory limit): the shim stages through pageable memory
    p = subprocess.run([exe], capture_output=True, text=True,
                       env=dict(os.environ, MOCK_PINNED_ALLOC_FAILS="1"))
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 failures" in p.stdout


@pytest.mark.gpu
def test_reference_unit_tests_through_shim(capi):
    _build()
    args = [BIN]
    try:
        capi.OctHandle(0.1).close()
        args.append("--octree")
    except capi.PgmoError:
        pass
    p = subprocess.run(args, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 failures" in p.stdout
