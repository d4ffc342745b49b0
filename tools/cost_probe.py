#!/usr/bin/env python
"""Time the tcgen05 cost construction (b200ot_cost) with the bf16 6-term split and the fp16 3-term split, and report
the error of both against float64 on a sample of rows.  CUDA events, median of 7."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch  # noqa: E402

from b200ot import ops  # noqa: E402


def med(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2]


def main():
    dev = torch.device("cuda", 0)
    out = []
    for n in [int(s) for s in (sys.argv[1] if len(sys.argv) > 1 else "4096,16384,65536").split(",")]:
        d = 512
        gen = torch.Generator(device="cpu").manual_seed(20251121)
        X = torch.randn(n, d, generator=gen)
        Y = torch.randn(n, d, generator=gen) + 0.5 * torch.randn(1, d, generator=gen)
        X = (X / X.norm(dim=1, keepdim=True)).to(dev)
        Y = (Y / Y.norm(dim=1, keepdim=True)).to(dev)
        Cm = ops.empty_matrix(n, n, dev)
        rows = torch.arange(0, n, max(1, n // 256), device=dev)[:256]
        ref = (X[rows].double() ** 2).sum(1)[:, None] + (Y.double() ** 2).sum(1)[None, :] - 2.0 * X[rows].double() @ Y.double().T
        rec = {"n": n, "d": d}
        only = sys.argv[2] if len(sys.argv) > 2 else None
        for terms in (6, "f16", "f16x4"):
            if only is not None and str(terms) != only:
                continue
            ms = med(lambda: ops.cost_matrix(X, Y, out=Cm, impl="tc", terms=terms))
            prod = {6: 6, "f16": 3, "f16x4": 4}[terms]
            rec[f"ms_{terms}"] = ms
            rec[f"executed_tflops_{terms}"] = 2.0 * n * n * d * prod / ms / 1e9
            rec[f"max_abs_err_{terms}"] = float((Cm[rows].double() - ref).abs().max())
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del Cm, ref
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", "cost_probe.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
