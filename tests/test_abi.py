"""The C ABI: the library builds, loads and exports exactly what include/b200ot.h declares,
and the ctypes table in b200ot/_lib.py matches it.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200ot.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ot_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native_lib):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(native_lib)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/b200ot.h but not exported: {missing}"


def test_ctypes_table_matches_header(native_lib):
    from b200ot import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.b200ot_version() == 100
    assert lib.b200ot_strerror(0) == b"ok"
    assert b"workspace" in lib.b200ot_strerror(-2)


def test_argument_counts_match_header(native_lib):
    from b200ot import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        count = 0 if params in ("", "void") else params.count(",") + 1
        assert count == len(args), (name, count, len(args))


def test_workspace_query_is_host_only(native_lib):
    from b200ot import _lib
    lib = _lib.load()
    small = lib.b200ot_sinkhorn_workspace_bytes(64, 64)
    big = lib.b200ot_sinkhorn_workspace_bytes(65536, 65536)
    assert 0 < small < big < 200 * 2 ** 20
    assert lib.b200ot_sinkhorn_workspace_bytes(0, 5) == 0


def test_no_oracle_or_cpu_fallback_in_product():
    """The product package must not import the oracle or carry a CPU path."""
    pkg = os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU or PyTorch fallback", ""), (f, "mentions oracle")


def test_sass_has_bulk_copy_and_cluster_barrier(native_lib):
    """The fused sweep must really use the TMA bulk-copy engine and cluster barriers."""
    out = subprocess.run(["cuobjdump", "-sass", native_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UBLKCP" in out.stdout
    assert "UCGABAR" in out.stdout
    assert "MUFU.EX2" in out.stdout
