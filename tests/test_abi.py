"""CPU: the C-ABI library loads and exports every symbol include/pgmo_b200.h declares; the
product fails loudly without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgmo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgmo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(capi):
    declared = _declared_symbols()
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (pgmo_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in pgmo_b200.h but not exported: {missing}"
    # the ctypes table covers the header too
    assert set(declared) == set(capi.SIGNATURES)


def test_no_cpu_fallback(capi):
    lib = capi.load()
    if lib.pgmo_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(capi.PgmoError) as e:
        capi.DeformHandle(0)
    assert e.value.code == capi.PGMO_ERR_NO_DEVICE


def test_library_is_sm100a_only(capi):
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "kimera_pgmo_b200")
    bad = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+\"(\.\./)*oracle|liboracle", text, re.M):
                    bad.append(f)
    inc = os.path.join(ROOT, "include")
    for base, _, files in os.walk(inc):
        for f in files:
            text = open(os.path.join(base, f), errors="ignore").read()
            if "oracle/" in text and "#include" in text and re.search(r"#include\s+\".*oracle", text):
                bad.append(f)
    assert not bad, bad
