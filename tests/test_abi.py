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


def _sass_by_kernel(native_lib):
    """{demangled-ish kernel name: SASS text} from cuobjdump -sass of the in-tree library."""
    import re
    out = subprocess.run(["cuobjdump", "-sass", native_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    parts = re.split(r"\n\s*Function : ", out.stdout)
    return {p.split("\n", 1)[0].strip(): p for p in parts[1:]}


SASS_OPS = ("UTCHMMA", "LDTM", "UBLKCP", "SYNCS", "UCGABAR", "MUFU.EX2", "HMMA", "DFMA")


def test_sass_opcode_table_proves_tcgen05_tma_and_cluster_paths(native_lib):
    """Per-kernel SASS opcode counts (also written to profiles/r02_sass_opcodes.json).  tcgen05.mma = UTCHMMA,
    tcgen05.ld = LDTM, TMA bulk copy = UBLKCP, mbarrier = SYNCS, cluster barrier = UCGABAR (B200_PROFILING.md)."""
    import json
    kernels = _sass_by_kernel(native_lib)
    table = {}
    for name, text in kernels.items():
        row = {op: text.count(op) for op in SASS_OPS}
        if any(row.values()):
            table[name] = row

    def find(sub):
        hits = [k for k in table if sub in k]
        assert hits, (sub, sorted(table))
        return {op: sum(table[k][op] for k in hits) for op in SASS_OPS}

    for kern in ("cost_tc_kernel", "apply_tc_kernel"):  # the two tensor-core kernels
        row = find(kern)
        assert row["UTCHMMA"] > 0 and row["LDTM"] > 0 and row["UBLKCP"] > 0 and row["SYNCS"] > 0, (kern, row)
        assert row["HMMA"] == row["UTCHMMA"], (kern, "legacy mma.sync found", row)  # HMMA only as a substring of UTCHMMA
    assert find("apply_tc_kernel")["MUFU.EX2"] > 0
    sweep = find("sweep_lite_kernel")
    assert sweep["UBLKCP"] > 0 and sweep["UCGABAR"] > 0 and sweep["MUFU.EX2"] > 0 and sweep["SYNCS"] > 0
    assert find("resident_kernel")["UBLKCP"] > 0
    assert find("sinkhorn_batched_kernel")["DFMA"] > 0  # float64 arithmetic of POT's loop
    out = os.path.join(ROOT, "profiles", "r02_sass_opcodes.json")
    with open(out, "w") as fh:
        json.dump({"source": "cuobjdump -sass b200ot/libb200ot.so (tests/test_abi.py)", "ops": list(SASS_OPS),
                   "kernels": {k: table[k] for k in sorted(table)}}, fh, indent=1)


def test_argument_validation_happens_before_any_device_work(native_lib):
    """Invalid calls return their status code without touching CUDA (so this runs on a CPU-only box): null
    pointers, non-positive sizes, shapes beyond what a kernel family covers.  No exceptions cross the ABI."""
    from b200ot import _lib
    lib = _lib.load()
    E_INVALID, E_UNSUPPORTED = -1, -4
    assert lib.b200ot_strerror(E_INVALID) == b"invalid argument"
    assert b"unsupported" in lib.b200ot_strerror(E_UNSUPPORTED)
    prm = _lib.Params(0.05, 10, 0.0, 10, 1, 0, 0, 0, 0)
    # Sinkhorn life cycle: missing workspace / marginals, empty problem
    assert lib.b200ot_sinkhorn_setup(4, 4, None, None, None, None, ctypes.byref(prm), None, 0, None) == E_INVALID
    assert lib.b200ot_sinkhorn_enqueue(None, 4, 4, 4, 1, 0, None, None) == E_INVALID
    assert lib.b200ot_sinkhorn_finish(0, 4, None, None, None, None, None, 0, None) == E_INVALID
    assert lib.b200ot_sinkhorn_describe(0, 4, ctypes.create_string_buffer(64), 64) == E_INVALID
    # batched / Gromov-Wasserstein one-CTA-per-problem kernels: size caps
    dummy = ctypes.c_void_p(256)  # never dereferenced: validation comes first
    assert lib.b200ot_sinkhorn_batched(dummy, None, None, 1, 129, 64, 0, dummy, dummy, ctypes.byref(prm), dummy, None,
                                       None, None, None, None) == E_UNSUPPORTED
    assert lib.b200ot_sinkhorn_batched(None, None, None, 1, 64, 64, 0, dummy, dummy, ctypes.byref(prm), dummy, None,
                                       None, None, None, None) == E_INVALID
    assert lib.b200ot_egw_batched(dummy, dummy, dummy, dummy, dummy, 1, 65, 10, 8, 8, 5e-3, 10, 5, 1e-3, 10, 10, 1e-3,
                                  dummy, dummy, dummy, None) == E_UNSUPPORTED
    assert lib.b200ot_egw_batched(dummy, dummy, dummy, dummy, dummy, 1, 10, 10, 8, 8, 0.0, 10, 5, 1e-3, 10, 10, 1e-3,
                                  dummy, dummy, dummy, None) == E_INVALID
    # peer exchange
    assert lib.b200ot_peer_exchange_bytes(0, 64) == 0 and lib.b200ot_peer_exchange_bytes(17, 64) == 0
    assert lib.b200ot_peer_exchange_bytes(8, 65536) == 2 * 8 * 65536 * 8
    assert lib.b200ot_peer_exchange_bytes(3, 100) == 2 * 3 * 128 * 8  # columns padded to 64
    assert lib.b200ot_peer_alloc(0, None, None) == E_INVALID
    assert lib.b200ot_peer_open(None, None) == E_INVALID
    assert lib.b200ot_sinkhorn_shard_finalize_peer(4, 4, None, None, 2, 1, 0, None) == E_INVALID


def test_host_api_rejects_cpu_tensors_and_bad_shapes(native_lib):
    """The Python operators fail loudly instead of falling back: CPU tensors, wrong dtypes and mismatched shapes
    raise B200OTError before any kernel is launched."""
    import torch
    from b200ot import ops, B200OTError
    x = torch.zeros(4, 8)
    with pytest.raises(B200OTError):
        ops.cost_matrix(x, x)
    with pytest.raises(B200OTError):
        ops.plan(torch.zeros(4, 4), torch.zeros(4), torch.zeros(4), 0.1)
    with pytest.raises(B200OTError):
        ops.egw_batched([], [])
    with pytest.raises(B200OTError):
        ops.egw_batched([x], [x, x])


def test_cost_split_codes_and_path_rule(native_lib):
    """Host logic without a device: the `terms` argument of the cost construction maps to the C-ABI codes of
    include/b200ot.h, and the streaming / online rule evaluates both per-iteration costs with the measured rates."""
    from b200ot import _lib
    from b200ot.online import MEASURED, choose_path
    src = open(HEADER).read()
    assert int(re.search(r"#define B200OT_TERMS_F16_3 (\d+)", src).group(1)) == _lib.TERMS_F16_3 == _lib.split_terms("f16")[0]
    assert int(re.search(r"#define B200OT_TERMS_F16_4 (\d+)", src).group(1)) == _lib.TERMS_F16_4 == _lib.split_terms("f16x4")[0]
    assert _lib.split_terms(6) == (6, 6, False) and _lib.split_terms("f16") == (19, 3, True)
    assert _lib.split_terms("f16x4")[1] == 4 and _lib.split_terms(1)[1] == 1
    with pytest.raises(ValueError):
        _lib.split_terms(5)
    # C fits: stream (4nm bytes per iteration beat 2nmd * products flops at the measured rates); C does not fit: online
    assert choose_path(65536, 65536, 512, 180 << 30) == "streaming"
    assert choose_path(300000, 300000, 512, 180 << 30) == "online"
    assert choose_path(65536, 65536, 8, 180 << 30) == "streaming"  # every panel is also written and read once
    fast_tensor = dict(MEASURED, online_tflops_executed=1e9)       # rule is driven by the rates, not hard-wired
    assert choose_path(65536, 65536, 512, 180 << 30, rates=fast_tensor) == "streaming"
    assert choose_path(65536, 65536, 512, 8 << 30) == "online"       # 16 GiB of C into 8 GiB of free memory
