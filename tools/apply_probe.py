#!/usr/bin/env python
"""Time the plan-application epilogues (tcgen05 single-pass kernel vs the generic SIMT kernel) with CUDA events.

    python tools/apply_probe.py [--sizes 4096,16384,65536] [--dv 512] [--out gpurun_out/apply_probe.json]

Reports ms per launch, executed/algorithmic TFLOP/s (2 n m dv algorithmic; x3 executed for the bf16 split) and
the bytes of C per second; inputs larger than L2 at >= 16384, an L2 flush between launches below that."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from b200ot import ops  # noqa: E402


def timeit(fn, reps, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,16384,65536")
    ap.add_argument("--dv", type=int, default=512)
    ap.add_argument("--simt-max", type=int, default=16384, help="largest n for which the SIMT kernel is also timed")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "apply_probe.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB > L2
    res = []
    for n in [int(s) for s in args.sizes.split(",")]:
        m, d = n, args.dv
        gen = torch.Generator(device="cpu").manual_seed(7)
        X = torch.randn(n, d, generator=gen)
        Y = torch.randn(m, d, generator=gen)
        X = (X / X.norm(dim=1, keepdim=True)).to(dev)
        Y = (Y / Y.norm(dim=1, keepdim=True)).to(dev)
        C = ops.cost_matrix(X, Y)
        a = torch.full((n,), 1.0 / n, device=dev)
        f, g, _ = ops.sinkhorn_potentials(C, a, a, 0.05, max_iter=10, tol=0.0)
        fl = flush if n * m * 4 < 4 * 126e6 else None
        reps = 5 if n >= 32768 else 11
        rec = {"n": n, "m": m, "dv": d}
        t = timeit(lambda: ops.apply_plan(C, f, g, 0.05, Y, normalise=True, impl="tc"), reps, fl)
        alg = 2.0 * n * m * d
        rec["tc_ms"] = t
        rec["tc_alg_tflops"] = alg / t / 1e9
        rec["tc_executed_tflops"] = 3 * alg / t / 1e9
        rec["tc_c_gbs"] = 4.0 * n * m / t / 1e6
        t = timeit(lambda: ops.apply_plan(C, f, g, 0.05, X, normalise=False, transpose=True, impl="tc"), reps, fl)
        rec["tc_transposed_ms"] = t
        t = timeit(lambda: ops.envelope_bwd(C, f, g, 0.05, X, Y, impl="tc"), reps, fl)
        rec["envelope_bwd_tc_ms"] = t
        rec["envelope_alg_tflops"] = 2 * alg / t / 1e9
        if n <= args.simt_max:
            rec["simt_ms"] = timeit(lambda: ops.apply_plan(C, f, g, 0.05, Y, normalise=True, impl="simt"), 3, fl)
            rec["envelope_bwd_simt_ms"] = timeit(lambda: ops.envelope_bwd(C, f, g, 0.05, X, Y, impl="simt"), 3, fl)
        print(json.dumps(rec), flush=True)
        res.append(rec)
        del C, X, Y
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
