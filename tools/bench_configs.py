#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (informational; bench.py is the contract benchmark).

  C1  64 x 64, d = 512, eps = 0.05, 200 iterations through the NumPy-facing drop-in (latency)
  C2  4096 independent 64 x 64 problems from embeddings, one CTA each, float64 (problems/s)
  C3  n = m = 4096 cohort: cost + 200 iterations + barycentric projection + fusion head fwd/bwd

    python tools/bench_configs.py > profiles/r01_configs.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]

import numpy as np
import torch

import b200ot
from b200ot import ops
from b200ot.fusion import OTFusionHead
from oracle import ot_oracle as orc


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    dev = torch.device("cuda", 0)
    out = {}
    eps = 0.05
    # ---- C1
    X, Y = orc.synthetic_embeddings(64, 64, 512, config_index=0)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(64) / 64
    t_gpu = timed(lambda: b200ot.sinkhorn(a, a, C, eps, numItermax=200, stopThr=0.0, warn=False), reps=10)
    t0 = time.perf_counter()
    for _ in range(5):
        orc.sinkhorn_knopp(a, a, M=C, reg=eps, numItermax=200, stopThr=0.0)
    t_cpu = (time.perf_counter() - t0) / 5
    out["C1_64x64_200it_numpy_api"] = {"b200_ms": 1e3 * t_gpu, "cpu_oracle_ms": 1e3 * t_cpu,
                                      "note": "host NumPy in / out, per-iteration kernel launches (2 per iteration)"}
    Xd, Yd = torch.tensor(X, device=dev)[None], torch.tensor(Y, device=dev)[None]
    ad = torch.tensor(a, dtype=torch.float32, device=dev)
    t_b1 = timed(lambda: ops.sinkhorn_batched(ad, ad, eps, X=Xd, Y=Yd, max_iter=200, tol=0.0), reps=10)
    out["C1_64x64_200it_batched_kernel_batch1"] = {"b200_ms": 1e3 * t_b1, "note": "one CTA, float64, whole solve in one launch"}
    # ---- C2
    B = 4096
    gen = torch.Generator(device="cpu").manual_seed(20251118 + 1)
    Xb = torch.randn(B, 64, 512, generator=gen)
    Yb = torch.randn(B, 64, 512, generator=gen) + 0.5 * torch.randn(B, 1, 512, generator=gen)
    Xb = (Xb / Xb.norm(dim=2, keepdim=True)).to(dev)
    Yb = (Yb / Yb.norm(dim=2, keepdim=True)).to(dev)
    t_c2 = timed(lambda: ops.sinkhorn_batched(ad, ad, eps, X=Xb, Y=Yb, max_iter=200, tol=0.0), reps=5)
    t_c2c = timed(lambda: ops.sinkhorn_batched(ad, ad, eps, X=Xb, Y=Yb, max_iter=2000, tol=1e-9), reps=5)
    P, lg = ops.sinkhorn_batched(ad, ad, eps, X=Xb, Y=Yb, max_iter=2000, tol=1e-9)
    out["C2_batched_4096x(64x64)_d512"] = {
        "fixed_200_iterations": {"ms": 1e3 * t_c2, "problems_per_s": B / t_c2},
        "pot_rule_to_convergence": {"ms": 1e3 * t_c2c, "problems_per_s": B / t_c2c,
                                    "mean_iterations": float(lg["n_iter"].float().mean())},
        "cpu_oracle_problems_per_s_200it": 1.0 / t_cpu,
        "hbm_bytes_per_problem": 4 * (64 + 64) * 512 + 4 * 64 * 64,
    }
    # ---- C3
    n = m = 4096
    X3, Y3 = orc.synthetic_embeddings(n, m, 512, config_index=2)
    x3, y3 = torch.tensor(X3, device=dev), torch.tensor(Y3, device=dev)
    a3 = torch.full((n,), 1.0 / n, device=dev)
    Cm = ops.cost_matrix(x3, y3)
    t_cost = timed(lambda: ops.cost_matrix(x3, y3, out=Cm))
    st = ops.SinkhornStepper(Cm, a3, a3, eps, max_iter=200, tol=0.0)

    def solve():
        st.reset()
        st.enqueue(200)
    t_solve = timed(solve)
    f, g, _ = st.finish()
    t_bary = timed(lambda: ops.apply_plan(Cm, f, g, eps, y3, normalise=True))
    head = OTFusionHead(512, 8, dropout=0.1).to(dev).train()
    Bsz = 32
    mri = torch.randn(Bsz, 512, device=dev)
    pet = torch.randn(Bsz, 512, device=dev, requires_grad=True)
    p2m = torch.randn(Bsz, 512, device=dev)
    mf = torch.randn(Bsz, 512, device=dev)
    T = torch.softmax(torch.randn(512, 512, device=dev), dim=1) / 512

    def fusion_step():
        attn, z, loss = head(mri, pet, p2m, mf, T, training=True)
        (attn.sum() + loss).backward()
    t_fuse = timed(fusion_step)
    out["C3_cohort_4096x4096_d512"] = {
        "cost_tcgen05_ms": 1e3 * t_cost, "sinkhorn_200it_ms": 1e3 * t_solve,
        "iterations_per_s": 200 / t_solve, "barycentric_projection_ms": 1e3 * t_bary,
        "fusion_head_fwd_bwd_batch32_ms": 1e3 * t_fuse,
        "note": "C = 64 MiB is L2-resident: launch- and latency-bound, not HBM-bound",
    }
    # ---- iteration rate by size: resident (persistent cooperative) kernel vs one launch per sweep
    sweep = {}
    for nn in (512, 1024, 2048, 4096, 8192):
        Xs, Ys = orc.synthetic_embeddings(nn, nn, 512, config_index=3)
        Cs = ops.cost_matrix(torch.tensor(Xs, device=dev), torch.tensor(Ys, device=dev))
        an = torch.full((nn,), 1.0 / nn, device=dev)
        rec = {}
        for label, env in (("resident", {"B200OT_RESIDENT": "1"}), ("resident_forward_only", {"B200OT_RESIDENT": "1", "B200OT_RES_SNAKE": "0"}),
                           ("resident_2x256_threads", {"B200OT_RESIDENT": "1", "B200OT_RES_WIDE": "0"}),
                           ("per_sweep_launches", {"B200OT_RESIDENT": "0"})):
            if label == "resident_forward_only" and nn < 4096:
                continue
            if label == "resident_2x256_threads" and nn <= 4096:
                continue
            os.environ.update(env)
            try:
                stp = ops.SinkhornStepper(Cs, an, an, eps, max_iter=200, tol=0.0)
                if label == "per_sweep_launches":
                    stp.build_graph(10)
                for _ in range(3):
                    stp.reset()
                    stp.run(200)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    stp.reset()
                    stp.run(200)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                _, _, inf = stp.finish()
                assert inf["n_iter"] == 200 and inf["status"] == 0, inf
                rec[label] = {"ms_200it": ms, "iterations_per_s": 200e3 / ms, "us_per_iteration": 5.0 * ms,
                              "matrix_GBps": 4.0 * nn * nn * 200 / (ms * 1e-3) / 1e9,
                              "kernel": ops.describe_kernel(nn, nn)}
            finally:
                for k in env:
                    os.environ.pop(k, None)
        sweep[f"n=m={nn}"] = rec
    out["sinkhorn_200it_by_size"] = sweep
    # ---- reference-native feature problems (MRI_PET_OT_nojax.py:91-145): d x d plan from 64 samples
    fot = {}
    for dd in (512, 2048):
        rng = np.random.default_rng(dd)
        Xf = rng.standard_normal((64, dd)).astype(np.float32)
        Yf = (rng.standard_normal((64, dd)) + 0.3).astype(np.float32)
        data = ({0: Xf}, {0: Yf})
        Ts = {0: np.eye(64) / 64}
        t_dev = timed(lambda: b200ot.get_feature_coupling_pot(data, Ts, eps=5e-3 if dd == 2048 else 1e-2), reps=5)
        t0 = time.perf_counter()
        orc.get_feature_coupling_pot(data, Ts, eps=5e-3 if dd == 2048 else 1e-2)
        t_ref = time.perf_counter() - t0
        fot[f"d={dd}"] = {"b200_ms_numpy_in_out": 1e3 * t_dev, "cpu_oracle_ms": 1e3 * t_ref}
    out["feature_coupling_pot"] = fot
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
