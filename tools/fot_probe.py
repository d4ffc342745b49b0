#!/usr/bin/env python
"""Time the FOT feature cost M = t1 (+) t2 - 2 X^T Ts Y (MRI_PET_OT_nojax.py:121-136) at the reference-native shapes:
tcgen05 chain (b200ot_fot_cost_tc) vs fp32 FMA kernel (b200ot_fot_cost), CUDA events, median of 21."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch  # noqa: E402

from b200ot import ops  # noqa: E402


def med(fn, reps=21):
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
    for n, d in ((64, 512), (128, 2048), (512, 2048), (2048, 4096)):
        g = torch.Generator(device="cpu").manual_seed(n + d)
        X = torch.randn(n, d, generator=g).to(dev)
        Y = torch.randn(n, d, generator=g).to(dev)
        Ts = torch.eye(n, device=dev) / n
        w = Ts.sum(1)
        rec = {"n": n, "d": d, "flops": 2.0 * d * d * n + 2.0 * n * n * d}
        for impl in ("tc", "simt"):
            rec[impl + "_ms"] = med(lambda: ops.fot_cost(X, Y, Ts, w, w, impl=impl))
        a = ops.fot_cost(X, Y, Ts, w, w, impl="tc")
        b = ops.fot_cost(X, Y, Ts, w, w, impl="simt")
        rec["max_abs_diff_over_max"] = float((a - b).abs().max() / b.abs().max())
        rec["tc_tflops"] = rec["flops"] / rec["tc_ms"] / 1e9
        print(json.dumps(rec), flush=True)
        out.append(rec)
    with open(os.path.join(ROOT, "gpurun_out", "fot_probe.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
