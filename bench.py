#!/usr/bin/env python
"""Headline benchmark: log-domain Sinkhorn iterations/s (and HBM GB/s) at n = m = 65536, d = 512.

    python bench.py --gpus 1 --steps K --warmup W            # B200 arm
    torchrun ... bench.py --gpus N ...                        # row-sharded, one rank per GPU
    python bench.py --impl reference ...                      # the reference's CPU algorithm

One *step* is one OT solve of the BASELINE workload: ITERS Sinkhorn iterations (g update, f update,
fused marginal check) over the cost matrix already resident in HBM.  `value` = iterations/s of the
whole job; `e2e` = the same metric through the public API with HOST (pinned) embeddings in and host
potentials out (H2D copy, cost construction, solve, D2H read inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "sinkhorn_iterations_per_s"
UNIT = "iterations/s"
EPS = 0.05
D = 512


def _traffic(kernel_desc, n_loc, m):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r02_traffic.json), when this run uses the captured configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            for rec in json.load(fh):
                if rec["n_local"] == n_loc and rec["m"] == m and rec["kernel"] in kernel_desc:
                    return rec["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (a thread in this process).
    An `nvidia-smi -lms` child polling the same GPU was measured to slow the rank it watches, and through the
    per-iteration all-reduce every other rank: 1.6k -> 2.1k it/s at 8 GPUs once it was replaced."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period_s=0.2):
        self.index, self.period = index, period_s
        self.sm, self.mask, self.max_sm = [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:  # NVML missing: say so instead of inventing numbers
            self.err = repr(e)

    def _sample(self):
        self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        try:
            self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception as e:
                self.err = repr(e)
                return
            self.stop_flag.wait(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        try:
            self._sample()  # at least one sample inside the region even for very short runs
        except Exception:
            pass
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = sorted(self.sm)
        reasons = [nm for bit, nm in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "source": "nvml"}


def synthetic_rows(n_total, m, lo, hi, seed, device):
    """Rows [lo, hi) of X and all of Y for the SURVEY 8(d) synthetic workload, generated on the host
    in blocks (torch CPU generator, so every rank sees the same Y) and L2-normalised."""
    import torch
    gen = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.randn(n_total, D, generator=gen)
    Y = torch.randn(m, D, generator=gen) + 0.5 * torch.randn(1, D, generator=gen)
    X = X[lo:hi]
    X = X / X.norm(dim=1, keepdim=True)
    Y = Y / Y.norm(dim=1, keepdim=True)
    return X.contiguous(), Y.contiguous()


# --------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref, staged by oracle/build_ref.py) -- bench-only use of oracle/
# --------------------------------------------------------------------------------------
def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed to run on."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pin_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference would use every core.  Must run before numpy loads."""
    nt = str(host_threads())
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = nt


class CpuArm:
    """`sinkhorn_scaling` of perturbot/perturbot/match/utils.py:6-115 -- the NumPy Sinkhorn-Knopp that
    MRI_PET_OT_nojax.py:143's ot.sinkhorn executes arithmetically -- on the leading n_s x n_s block of the bench
    workload (same generator, same rows), float64, K = exp(-C/eps) prepared outside the timed region.  One step =
    one call with numItermax = iters (stopThr = 0), so the call's own set-up (Kp = K / a) and return value
    (diag(u) K diag(v)) are inside the time, amortised over `iters` iterations instead of a solve's 200."""

    def __init__(self, n_s, n_full):
        import numpy as np
        from oracle import build_ref
        from oracle import ot_oracle as orc
        self.np = np
        self.ref = build_ref.load()
        self.kind = "reference" if self.ref is not None else "port"
        self.orc = orc
        self.n_s, self.n_full = n_s, n_full
        X, Y = synthetic_rows(n_full, n_full, 0, n_s, 20251118 + 3, None)
        K = orc.sqeuclid_cost(X.numpy(), Y[:n_s].numpy())
        K *= -1.0 / EPS
        np.exp(K, out=K)
        self.K = K
        self.a = np.full(n_s, 1.0 / n_s)
        try:
            from threadpoolctl import threadpool_info
            self.blas = [f"{d.get('internal_api')}:{d.get('num_threads')}" for d in threadpool_info()]
        except Exception:
            self.blas = []

    def step(self, iters):
        t0 = time.perf_counter()
        if self.ref is not None:
            self.ref.sinkhorn_scaling(self.a, self.a, self.K, numItermax=iters, stopThr=0.0)
        else:
            self.orc.sinkhorn_knopp(self.a, self.a, K=self.K, numItermax=iters, stopThr=0.0, err_norm="l2sq")
        return time.perf_counter() - t0

    def what(self):
        return ("sinkhorn_scaling of the reference (perturbot/perturbot/match/utils.py:6-115, staged unmodified in "
                "oracle/_ref)" if self.ref is not None else
                "oracle port of sinkhorn_scaling (oracle/_ref not staged)")


def cpu_sizes(requested):
    """Largest block the host can hold: the reference needs K, Kp and two n^2 temporaries in float64 (32 n^2 B)."""
    if requested:
        return requested
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    return 32768 if avail > 48 * 2**30 else 16384


def cpu_measure(args, steps, warmup, n_big):
    """`warmup` short calls and `steps` timed calls at n = 8192, then ONE call each at n = 16384 and (when the host
    has the memory) n = 32768.  The rate at the largest block is extrapolated to the 65536^2 problem by the O(n^2)
    cost per iteration (x (n_big/n)^2) -- K and Kp at 65536^2 need 69 GB in float64 -- and the
    smaller blocks show that the scaling holds."""
    n_full = args.n
    n_small = min(8192, n_full)
    arm = CpuArm(n_small, n_full)
    for _ in range(warmup):
        arm.step(2)
    times = [arm.step(args.cpu_iters) for _ in range(steps)]
    rate_small = args.cpu_iters / (sum(times) / len(times))
    kind, what, blas = arm.kind, arm.what(), arm.blas
    del arm
    rates = [(n_small, rate_small, sum(times) / len(times))]
    for nb in (16384, 32768):
        if nb <= n_small or nb > n_big or nb > n_full:
            continue
        big = CpuArm(nb, n_full)
        big.step(1)  # touch the pages once
        t = big.step(args.cpu_iters)
        rates.append((nb, args.cpu_iters / t, t))
        del big
    n_used, rate_used, _ = rates[-1]
    scaled = rate_used * (n_used / n_full) ** 2
    per_size = "; ".join(f"n=m={nn}: {rr:.3f} it/s ({tt:.1f} s per call, x(n/{n_full})^2 -> {rr * (nn / n_full) ** 2:.4f})"
                         for nn, rr, tt in rates)
    sample = (f"{what}; float64; calls of {args.cpu_iters} iterations on the leading n x n block of the n=m={n_full} "
              f"workload, {steps} timed calls at n={n_small} then one call per larger block: {per_size}; value = the "
              f"n={n_used} rate x ({n_used}/{n_full})^2; BLAS threads {blas}")
    return scaled, kind, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    scaled, kind, sample = cpu_measure(args, args.steps, args.warmup, cpu_sizes(args.cpu_sample))
    ms = 1e3 * args.iters / scaled
    line = {"impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"log-domain/kernel-domain Sinkhorn n=m={args.n} d={D} eps={EPS}, "
                                   f"{args.iters} iterations per solve (BASELINE configs[3])", "n": args.n, "m": args.n,
                       "d": D, "eps": EPS, "iterations_per_step": args.iters},
            "cpu_baseline": {"value": scaled, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
PARITY_ROWS = 1024


def _event_ms(torch, fn, reps, warm=2):
    for _ in range(warm):
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
    ts.sort()
    return ts[len(ts) // 2]


def extras(torch, dist, ops, world, rank, dev, args):
    """The other BASELINE configs, measured in the same run so that the driver sees them:
    C2  4096 independent 64 x 64 problems from embeddings (d = 512), float64 POT arithmetic, 200 iterations each,
        the batch split over the ranks with no communication -> problems/s of the whole job;
    C3  n = m = 4096 cohort: iterations/s of the resident kernel (C lives in L2), and the end-to-end time of
        pinned embeddings -> cost -> 200 iterations -> fused embedding -> envelope backward -> fused embedding D2H;
    online  the cost-free solver at n = m = 65536 (C rebuilt on tcgen05 every iteration, never in HBM).
    C3 and online run on rank 0 at N = 1 only (they do not shard; "replicas only")."""
    import b200ot
    out = {}
    # ---- C2: batched minibatch OT, batch split across the ranks
    B, nb, d = 4096, 64, D
    B_loc = B // world + (1 if rank < B % world else 0)
    gen = torch.Generator(device="cpu").manual_seed(20251118 + 1 + 7919 * rank)
    Xb = torch.randn(B_loc, nb, d, generator=gen)
    Yb = torch.randn(B_loc, nb, d, generator=gen) + 0.5 * torch.randn(B_loc, 1, d, generator=gen)
    Xb = (Xb / Xb.norm(dim=2, keepdim=True)).to(dev)
    Yb = (Yb / Yb.norm(dim=2, keepdim=True)).to(dev)
    ab = torch.full((nb,), 1.0 / nb, device=dev)
    its = 200
    holder = {}

    def run_c2():
        holder["r"] = ops.sinkhorn_batched(ab, ab, EPS, X=Xb, Y=Yb, max_iter=its, tol=0.0)
    if world > 1:
        dist.barrier()
    ms = _event_ms(torch, run_c2, reps=5)
    n_it = holder["r"][1]["n_iter"]
    assert int(n_it.min()) == its and int(n_it.max()) == its
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    out["c2_batched"] = {
        "workload": f"{B} independent {nb}x{nb} problems, d={d}, eps={EPS}, {its} iterations each (BASELINE configs[1]), "
                    f"float64 POT arithmetic, Gibbs kernel in registers, one problem per CTA, two CTAs per SM",
        "problems_per_s": B / (ms * 1e-3), "ms": ms, "n_gpus": world, "problems_per_gpu": B_loc,
        "split": "batch dimension over the ranks, no communication",
        "problem_iterations_per_s": B * its / (ms * 1e-3),
        "sm_ns_per_problem_iteration": ms * 1e6 * sm * world / (B * its),
        "hbm_bytes_per_problem": 4 * (2 * nb * d + nb * nb)}
    del Xb, Yb, holder
    if rank != 0 or world != 1:
        return out
    # ---- C3: cohort problem, n = m = 4096
    n3 = 4096
    X3h, Y3h = synthetic_rows(n3, n3, 0, n3, 20251118 + 2, None)
    X3p, Y3p = X3h.pin_memory(), Y3h.pin_memory()
    x3, y3 = X3p.to(dev), Y3p.to(dev)
    C3 = ops.cost_matrix(x3, y3)
    a3 = torch.full((n3,), 1.0 / n3, device=dev)
    st3 = ops.SinkhornStepper(C3, a3, a3, EPS, max_iter=its, tol=0.0)

    def run_c3():
        st3.reset()
        st3.run(its)
    ms3 = _event_ms(torch, run_c3, reps=5)
    f3, g3, inf3 = st3.finish()
    assert inf3["n_iter"] == its
    ms_apply = _event_ms(torch, lambda: ops.apply_plan(C3, f3, g3, EPS, y3, normalise=True), reps=11)
    ms_bwd = _event_ms(torch, lambda: ops.envelope_bwd(C3, f3, g3, EPS, x3, y3), reps=11)
    fused3 = torch.empty((n3, D), dtype=torch.float32).pin_memory()

    def e2e_c3():
        o = b200ot.sinkhorn_from_embeddings(X3p, Y3p, reg=EPS, numItermax=its, stopThr=0.0, V=Y3p, fused_out=fused3)
        xd, yd = X3p.to(dev, non_blocking=True), Y3p.to(dev, non_blocking=True)
        Cx = ops.cost_matrix(xd, yd)
        ff = torch.as_tensor(o["f"]).to(dev)
        gg = torch.as_tensor(o["g"]).to(dev)
        dx, dy = ops.envelope_bwd(Cx, ff, gg, EPS, xd, yd)
        dx.cpu()
    e2e_c3()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_c3()
    torch.cuda.synchronize()
    e2e3 = (time.perf_counter() - t0) / 3
    out["c3_cohort"] = {
        "workload": f"n=m={n3}, d={D}, eps={EPS}, {its} iterations (BASELINE configs[2]); C = 64 MiB lives in L2",
        "kernel": ops.describe_kernel(n3, n3), "iterations_per_s": its / (ms3 * 1e-3), "us_per_iteration": 1e3 * ms3 / its,
        "l2_gbs": 4.0 * n3 * n3 * its / (ms3 * 1e-3) / 1e9,
        "fused_embedding_ms": ms_apply, "envelope_backward_ms": ms_bwd,
        "fused_embedding_alg_tflops": 2.0 * n3 * n3 * D / ms_apply / 1e9,
        "e2e_ms": 1e3 * e2e3,
        "e2e_note": "pinned embeddings -> H2D -> cost -> 200 iterations -> fused embedding (tcgen05) -> D2H, then "
                    "cost + envelope backward (one tcgen05 launch for dX and dY) -> dX D2H"}
    del C3, st3
    # ---- C5: OT share of the end-to-end training step (reference CPU path of the OT leg vs device path)
    if not args.no_step:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_step
            out["c5_step"] = bench_step.measure(32, 96, steps=3, warm=2, dev=dev)
        except Exception as exc:  # noqa: BLE001 -- informational: never lose the headline line over it
            out["c5_step"] = {"error": repr(exc)}
        torch.cuda.empty_cache()
    # ---- online (cost-free) solver at the headline shape
    if not args.no_online:
        try:
            _online_extra(torch, ops, dev, args, out)
        except Exception as exc:  # noqa: BLE001
            out["online_c4"] = {"error": repr(exc)}
    return out


def _online_extra(torch, ops, dev, args, out):
    if True:
        from b200ot.online import OnlineSinkhorn
        n = m = args.n
        Xh, Yh = synthetic_rows(n, m, 0, n, 20251118 + 3, None)
        xo, yo = Xh.to(dev), Yh.to(dev)
        ao = torch.full((n,), 1.0 / n, device=dev)
        bo = torch.full((m,), 1.0 / m, device=dev)
        k_it = 4
        sol = OnlineSinkhorn(xo, yo, ao, bo, EPS, max_iter=10 ** 6, tol=0.0)
        sol.start()
        sol.run(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sol.run(k_it)
        e1.record()
        torch.cuda.synchronize()
        ms_o = e0.elapsed_time(e1) / k_it
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                tf_peak = float(json.load(fh)["bf16_tflops_sustained"])
        except Exception:
            tf_peak = 1400.0
        executed = sol.tensor_flops_per_iteration / ms_o / 1e9
        out["online_c4"] = {
            "workload": f"cost-free Sinkhorn n=m={n} d={D}: C rebuilt per iteration on tcgen05 in {sol.panel_rows}-row panels "
                        f"({sol.panel.numel() * 4 >> 20} MiB scratch, {sol.products}-product "
                        f"{'fp16' if sol.products in (3, 4) and sol.terms >= 19 else 'bf16'} split), consumed by the single-sweep "
                        f"kernel; the n x m matrix is never materialised",
            "iterations_per_s": 1e3 / ms_o, "ms_per_iteration": ms_o,
            "tensor_tflops_executed": executed, "tensor_tflops_algorithmic": executed / sol.products,
            "tensor_frac_of_sustained_bf16_peak": executed / tf_peak, "tensor_peak_tflops": tf_peak,
            "streaming_is_faster_by": None}
    return out




def parity_check(torch, dist, ops, world, rank, dev, Xh, Yh, Cmat, a_loc, f, g, n, m, info):
    """Values of the TIMED solve, checked after the timed region (never inside it).
    (1) g is replicated: every rank must hold the same bits (max and min over ranks of the int32 view agree).
    (2) rank 0 redoes rows [0, 1024) in float64 on the host (oracle.rows_given_g: cost rows from the embeddings, the
        f update and the plan rows that follow from the converged g) and compares the GPU's f and plan rows with it:
        max-normalised error, elementwise relative error on entries >= 1e-6 * max, |df| in units of eps.
    (3) the column marginals of the plan after the last f update are recomputed by an independent kernel
        (apply_plan_t on a column of ones, all-reduced over the row shards) and compared with b in L1."""
    out = {"g_bit_equal_across_ranks": True, "ranks": world}
    gi = g.view(torch.int32)
    if world > 1:
        hi, lo = gi.clone(), gi.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        out["g_bit_equal_across_ranks"] = bool(torch.equal(hi, lo))
    n_loc = Cmat.shape[0]
    cols = ops.apply_plan(Cmat, f, g, EPS, torch.ones((n_loc, 1), device=dev), transpose=True).reshape(-1).double()
    if world > 1:
        dist.all_reduce(cols)
    out["col_marginal_l1_recomputed"] = float((cols - 1.0 / m).abs().sum())
    out["n_iter"] = info["n_iter"]
    if rank != 0:
        return None
    import hashlib
    import numpy as np
    from oracle import ot_oracle as orc
    R = min(PARITY_ROWS, n_loc)
    g64 = g.double().cpu().numpy()
    C_rows = orc.sqeuclid_cost(Xh[:R].numpy(), Yh.numpy())
    f_ref, P_ref = orc.rows_given_g(C_rows, np.full(R, 1.0 / n), g64, EPS)
    P_gpu = ops.plan(Cmat[:R], f[:R], g, EPS).double().cpu().numpy()
    diff = np.abs(P_gpu - P_ref)
    mx = float(P_ref.max())
    mask = P_ref >= 1e-6 * mx
    out.update({
        "rows_checked": R,
        "plan_max_norm_err": float(diff.max() / mx),
        "plan_elementwise_rel_err_ge_1e-6max": float((diff[mask] / P_ref[mask]).max()),
        "entries_ge_1e-6max": int(mask.sum()),
        "f_abs_err_over_eps": float(np.abs(f[:R].double().cpu().numpy() - f_ref).max() / EPS),
        "cost_rows_abs_err": float(np.abs(Cmat[:R].double().cpu().numpy() - C_rows).max()),
        "tolerance": 1e-4,
        "g_sha256_16": hashlib.sha256(g.cpu().numpy().tobytes()).hexdigest()[:16],
        "oracle": "oracle.rows_given_g (float64) on rows [0, %d) with the GPU's converged g" % R,
    })
    out["ok"] = bool(out["g_bit_equal_across_ranks"] and out["plan_max_norm_err"] < 1e-4)
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from b200ot import ops, sharded
    import b200ot

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: b200ot has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = m = args.n
    iters = args.iters
    lo, hi = sharded.row_range(n, world, rank)
    seed = 20251118 + 3
    Xh, Yh = synthetic_rows(n, m, lo, hi, seed, dev)
    Xp, Yp = Xh.pin_memory(), Yh.pin_memory()
    n_loc = hi - lo
    a_loc = torch.full((n_loc,), 1.0 / n, dtype=torch.float32, device=dev)
    b = torch.full((m,), 1.0 / m, dtype=torch.float32, device=dev)
    Cmat = ops.cost_matrix(Xp.to(dev), Yp.to(dev))
    torch.cuda.synchronize()

    prm = ops.make_params(EPS, iters, 0.0, 10, 1, "l2", False, args.path)
    kern = sharded.CudaShardKernels(Cmat, a_loc, b, prm, path=args.path)
    balance = None
    if world > 1 and not args.no_balance:
        # Setup, outside the timed region: every iteration ends in an exchange of all ranks, so the loop runs at the
        # pace of the slowest GPU.  Time the local sweep of every rank (all ranks at once, clocks settled), split the
        # rows in proportion to the measured rates and rebuild this rank's rows of C when the split moves by > 2 %
        # (boxes with a 7 % spread between GPUs were seen; on a box with 1.8 % the repartition gained nothing:
        # profiles/r02_bench_n8_final.json).
        kern.setup()
        kern.finalize(kern.prologue(), True)
        dist.barrier()
        rate = sharded.measure_sweep_rate(kern)
        rates = [torch.zeros(1, device=dev) for _ in range(world)]
        dist.all_gather(rates, torch.tensor([rate], device=dev))
        rates = [float(r.item()) for r in rates]
        bounds = sharded.balanced_bounds(n, rates)
        moved = max(abs((bh - bl) - (sharded.row_range(n, world, r)[1] - sharded.row_range(n, world, r)[0]))
                    for r, (bl, bh) in enumerate(bounds)) / (n / world)
        balance = {"rows_per_ms_by_rank": [round(r, 1) for r in rates], "rows_by_rank": [bh - bl for bl, bh in bounds],
                   "applied": bool(moved > 0.02),
                   "note": "rows proportional to the measured local sweep rate of each GPU (setup, not timed)"}
        if balance["applied"]:
            del kern, Cmat
            torch.cuda.empty_cache()
            lo, hi = bounds[rank]
            Xh, Yh = synthetic_rows(n, m, lo, hi, seed, dev)
            Xp, Yp = Xh.pin_memory(), Yh.pin_memory()
            n_loc = hi - lo
            a_loc = torch.full((n_loc,), 1.0 / n, dtype=torch.float32, device=dev)
            Cmat = ops.cost_matrix(Xp.to(dev), Yp.to(dev))
            torch.cuda.synchronize()
            kern = sharded.CudaShardKernels(Cmat, a_loc, b, prm, path=args.path)
    comm = sharded.NcclComm() if (world > 1 and args.loop == "c") else None
    peer = None
    if world > 1 and args.loop == "peer":
        try:  # exchange buffers mapped into every rank with CUDA IPC; all ranks must agree on the outcome
            peer = sharded.PeerExchange(m)
            ok = 1
        except Exception as exc:  # noqa: BLE001
            peer, ok = None, 0
            if rank == 0:
                print(f"peer exchange unavailable ({exc}); using the NCCL loop", file=sys.stderr)
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if peer is not None:
                peer.close()
            peer, args.loop = None, "c"
            comm = sharded.NcclComm()
    solver = sharded.ShardedSinkhorn(kern, comm=comm, peer=peer)
    stepper = None
    resident = world == 1 and "resident_kernel" in ops.describe_kernel(n_loc, m)
    if resident:
        args.graph = 0  # short-iteration sizes: the whole solve is one persistent launch (csrc/resident.cu)
    if world == 1:
        stepper = ops.SinkhornStepper(Cmat, a_loc, b, EPS, max_iter=iters, tol=0.0, path=args.path)
        if args.graph:
            stepper.build_graph(args.graph)
    elif args.graph and comm is None and args.loop == "graph":
        solver.build_graph(args.graph)

    def step_device():
        if world == 1:
            stepper.reset()
            stepper.run(iters)
        else:
            solver.start()
            solver.run(iters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # correctness guard on the timed state: finished all iterations, finite error
    if world == 1:
        f_fin, g_fin, info = stepper.finish()
    else:
        f_fin, g_fin, info = kern.finish()
    diag = bool(os.environ.get("B200OT_FUSED_MODE"))  # diagnostic modes do not run the real arithmetic
    if not diag:
        assert info["n_iter"] == iters and info["status"] == 0, info
    parity = None
    if not diag and not args.no_parity:
        parity = parity_check(torch, dist, ops, world, rank, dev, Xh, Yh, Cmat, a_loc, f_fin, g_fin, n, m, info)
    ms_per_step = ms / args.steps
    value = iters / (ms_per_step * 1e-3)

    # ---- end to end through the public API: pinned host embeddings -> potentials AND the fused embedding
    # (barycentric projection of the PET embeddings through the plan, single-pass tcgen05 kernel) on the host
    fused_host = torch.empty((n_loc, D), dtype=torch.float32).pin_memory()
    fg_host = torch.empty(n_loc + m, dtype=torch.float32).pin_memory()

    def step_e2e():
        if world == 1:
            out = b200ot.sinkhorn_from_embeddings(Xp, Yp, reg=EPS, numItermax=iters, stopThr=0.0, path=args.path,
                                                  V=Yp, fused_out=fused_host)
            return out["err"]
        xd = Xp.to(dev, non_blocking=True)
        yd = Yp.to(dev, non_blocking=True)
        ops.cost_matrix(xd, yd, out=Cmat)
        f, g, inf = sharded.solve_sharded(Cmat, a_loc, b, EPS, max_iter=iters, tol=0.0, path=args.path, comm=comm,
                                          peer=peer)
        fused = ops.apply_plan(Cmat, f, g, EPS, yd, normalise=True)  # rows are local: no communication
        fused_host.copy_(fused, non_blocking=True)
        fg_host[:n_loc].copy_(f, non_blocking=True)
        fg_host[n_loc:].copy_(g, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return inf["err"]

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = iters / e2e_s
    h2d = (Xp.numel() + Yp.numel()) * 4
    d2h = (n_loc + m) * 4 + 32 + n_loc * D * 4  # potentials, result block, fused embedding (n_loc x 512 fp32)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # bounded sample (~20 s): calls at n = 8192 and one at 16384; `--impl reference` also measures n = 32768
        scaled, kind, sample = cpu_measure(args, 2, 1, 16384)
        cpu_baseline = {"value": scaled, "unit": UNIT, "cores": host_threads(), "kind": kind, "sample": sample}
    extra = None
    if not diag and not args.no_extras:
        stepper = None  # release the workspace
        try:
            extra = extras(torch, dist, ops, world, rank, dev, args)
            if extra.get("online_c4") and "iterations_per_s" in extra["online_c4"]:
                extra["online_c4"]["streaming_is_faster_by"] = value / extra["online_c4"]["iterations_per_s"]
        except Exception as exc:  # noqa: BLE001 -- informational block: never lose the headline line over it
            extra = {"error": repr(exc)}
    if rank != 0:
        if world > 1:
            _shutdown(dist)
        return
    peak, peak_src = _peaks()
    alg_bytes = 4.0 * (n / world) * m  # one fp32 read of C per iteration, per GPU (mean rows per rank)
    per_iter_s = ms_per_step * 1e-3 / iters
    achieved = alg_bytes / per_iter_s / 1e9
    # our kernels per step: init (init_state, init, colpass, finalize) + snapshot per enqueue + 2 per iteration;
    # sharded: setup (2) + prologue (colpass, reduce_parts) + finalize, then sweep + reduce_parts + finalize per iteration
    from b200ot import _lib as _b200ot_lib
    lib_counter = _b200ot_lib.load().b200ot_sinkhorn_counter
    n_enq = (iters // args.graph + (1 if iters % args.graph else 0)) if args.graph else 1
    fused_form = lib_counter(0) > 0 and lib_counter(1) == 0 and (world == 1 or args.loop == "peer")
    if fused_form:
        # persistent fused kernel: N = 1: init (4) + per graph replay / enqueue a snapshot and ONE launch that runs its
        # iterations; N > 1: setup (2) + first g update (3) + one launch per host window of 10 iterations
        win = int(os.environ.get("B200OT_SHARD_WINDOW", "10"))
        launches_per_step = (4 + 2 * n_enq) if world == 1 else (5 + (iters + win - 1) // win)
    else:
        # N = 1: sweep + finalize; peer loop: sweep + ONE fold / push / poll / finalize launch (peer_tail_kernel);
        # NCCL / Python loops and B200OT_PEER_TAIL=0: sweep, reduce(+push), finalize
        merged_tail = args.loop == "peer" and os.environ.get("B200OT_PEER_TAIL", "1") != "0"
        launches_per_step = (4 + n_enq + 2 * iters) if world == 1 else (5 + (2 if merged_tail else 3) * iters)
    if resident:
        launches_per_step = 4 + 1 + 1  # init, snapshot, one resident launch for all iterations
    c_bytes = 4.0 * n_loc * m
    l2_note = ("cost matrix (%.1f GiB per GPU) is far larger than L2, no flush needed" % (c_bytes / 2**30)
               if c_bytes > 4 * 126e6 else
               "cost matrix (%.0f MiB) is comparable to / smaller than the 126 MB L2 and is re-read by every iteration "
               "of a solve by design; not flushed between iterations (a solve is the timed unit)" % (c_bytes / 2**20))
    kernel_desc = ops.describe_kernel(n_loc, m)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"log-domain Sinkhorn n=m={n} d={D} eps={EPS}, {iters} iterations per solve "
                               f"(BASELINE configs[3]; single-sweep fused kernel, C resident in HBM)",
                   "n": n, "m": m, "d": D, "eps": EPS, "iterations_per_step": iters, "path": args.path,
                   "rows_per_gpu": n_loc if balance is None or not balance["applied"] else balance["rows_by_rank"],
                   "balance": balance, "kernel": kernel_desc,
                   "launch": ("one persistent launch per solve" if resident else
                              (f"CUDA graph, {args.graph} iterations per replay" if args.graph else "eager launches") +
                              (", each replay = snapshot + ONE persistent cooperative cluster launch (sweep, fold, finalize and "
                               "stopping rule of all its iterations)" if fused_form else ""))
                   if world == 1 else f"loop={args.loop}", "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                   "l2": l2_note},
        "hbm_gbs": achieved * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _traffic(kernel_desc, n_loc, m),
                     "traffic_source": "static: dram__bytes_read+write per launch from the committed ncu --set full capture "
                                       "of this kernel configuration (profiles/r02_traffic.json); not re-measured in this run",
                     "peak_source": peak_src,
                     "note": "achieved = 4*n_local*m bytes per iteration / (step time / iterations); the step time "
                             "includes the finalize kernel and, for N>1, the exchange of the column sums (peer-memory push or NCCL)"},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "pinned host embeddings -> H2D -> cost construction (tcgen05) -> solve -> fused embedding "
                        "diag(1/P1) P Y (tcgen05, one pass over C) -> potentials and fused embedding D2H (pinned)"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "parity": parity,
        "fused_iteration_launches": int(lib_counter(0)), "fused_iteration_fallbacks": int(lib_counter(1)),
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        _shutdown(dist)


def _shutdown(dist):
    """Leave without tearing NCCL down rank by rank (a captured graph or a second communicator can make
    destroy_process_group block); the line is already printed and flushed."""
    try:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=65536, help="problem side (n = m)")
    ap.add_argument("--iters", type=int, default=200, help="Sinkhorn iterations per step (one solve)")
    ap.add_argument("--path", default="auto", choices=["auto", "fused", "robust"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--graph", type=int, default=10, help="iterations per CUDA-graph replay at N=1 (0 = eager launches)")
    ap.add_argument("--loop", default="peer", choices=["peer", "c", "python", "graph"],
                    help="N>1: how the column sums are exchanged (peer = pushed as tagged words into peer memory over "
                         "NVLink by the kernels themselves, no collective call; c = ncclAllReduce queued from C on the "
                         "compute stream; python = torch.distributed all_reduce per iteration; graph = python loop captured)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="--impl reference: side of the largest block timed on the CPU (0 = 32768 if the host has "
                         "the memory for it, else 16384)")
    ap.add_argument("--cpu-iters", type=int, default=10, help="reference iterations per CPU call")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run value checks")
    ap.add_argument("--no-extras", action="store_true", help="skip the C2 / C3 / online measurements")
    ap.add_argument("--no-online", action="store_true", help="skip the online (cost-free) measurement")
    ap.add_argument("--no-step", action="store_true", help="skip the C5 end-to-end training-step measurement")
    ap.add_argument("--no-sampler", action="store_true", help="diagnostic: do not sample clocks")
    ap.add_argument("--no-balance", action="store_true",
                    help="N > 1: keep the even row split instead of rows proportional to the measured sweep rate of each GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        pin_host_threads()  # before numpy / BLAS load: torchrun exports OMP_NUM_THREADS=1
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
