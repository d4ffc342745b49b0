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
    (profiles/r01_traffic.json), when this run uses the captured configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
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
# CPU arms (the oracle's restatement of the reference algorithm; bench-only use of oracle/)
# --------------------------------------------------------------------------------------
def cpu_reference_rate(n_sample, iters, n_full, repeats=1):
    """Reference algorithm (kernel-domain Sinkhorn-Knopp in float64, MRI_PET_OT_nojax.py:143 /
    perturbot/match/utils.py:6-115) on an n_sample^2 slice of the workload; rate scaled to the full
    n_full^2 problem by the O(n^2) cost per iteration."""
    import numpy as np
    from oracle import ot_oracle as orc
    X, Y = orc.synthetic_embeddings(n_sample, n_sample, D, config_index=3)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n_sample) / n_sample
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.sinkhorn_knopp(a, a, M=C, reg=EPS, numItermax=iters, stopThr=0.0)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    rate_sample = iters / best
    return rate_sample * (n_sample / n_full) ** 2, rate_sample, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    try:
        import torch
        cores = torch.get_num_threads()
    except Exception:
        cores = os.cpu_count()
    n_full = args.n
    n_s = args.cpu_sample
    times = []
    for i in range(args.warmup + args.steps):
        scaled, raw, dt = cpu_reference_rate(n_s, args.cpu_iters, n_full)
        if i >= args.warmup:
            times.append((scaled, raw, dt))
    scaled = sum(t[0] for t in times) / len(times)
    raw = sum(t[1] for t in times) / len(times)
    ms = 1e3 * args.iters / scaled
    sample = (f"n=m={n_s} slice of the n=m={n_full} workload, {args.cpu_iters} iterations per step, float64 "
              f"kernel-domain Sinkhorn-Knopp ({raw:.2f} it/s on the slice), scaled by ({n_s}/{n_full})^2")
    line = {"impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"log-domain/kernel-domain Sinkhorn n=m={n_full} d={D} eps={EPS}, "
                                   f"{args.iters} iterations per solve", "n": n_full, "m": n_full, "d": D,
                       "eps": EPS},
            "cpu_baseline": {"value": scaled, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
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
        _, _, info = stepper.finish()
    else:
        _, _, info = kern.finish()
    if not os.environ.get("B200OT_FUSED_MODE"):  # diagnostic modes do not run the real arithmetic
        assert info["n_iter"] == iters and info["status"] == 0, info
    ms_per_step = ms / args.steps
    value = iters / (ms_per_step * 1e-3)

    # ---- end to end through the public API: pinned host embeddings -> potentials on the host
    def step_e2e():
        if world == 1:
            out = b200ot.sinkhorn_from_embeddings(Xp, Yp, reg=EPS, numItermax=iters, stopThr=0.0, path=args.path)
            return out["err"]
        xd = Xp.to(dev, non_blocking=True)
        yd = Yp.to(dev, non_blocking=True)
        ops.cost_matrix(xd, yd, out=Cmat)
        f, g, inf = sharded.solve_sharded(Cmat, a_loc, b, EPS, max_iter=iters, tol=0.0, path=args.path, comm=comm,
                                          peer=peer)
        f.cpu(), g.cpu()
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
    d2h = (n_loc + m) * 4 + 32

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        scaled, raw, dt = cpu_reference_rate(args.cpu_sample, args.cpu_iters, n)
        cpu_baseline = {"value": scaled, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"n=m={args.cpu_sample} slice, {args.cpu_iters} iterations, float64 kernel-domain "
                                  f"Sinkhorn-Knopp of the oracle ({raw:.2f} it/s, {dt:.1f} s), scaled by "
                                  f"({args.cpu_sample}/{n})^2 to the full problem"}
    if rank != 0:
        if world > 1:
            _shutdown(dist)
        return
    peak, peak_src = _peaks()
    alg_bytes = 4.0 * n_loc * m  # one fp32 read of this rank's rows of C per iteration
    per_iter_s = ms_per_step * 1e-3 / iters
    achieved = alg_bytes / per_iter_s / 1e9
    # our kernels per step: init (init_state, init, colpass, finalize) + snapshot per enqueue + 2 per iteration;
    # sharded: setup (2) + prologue (colpass, reduce_parts) + finalize, then sweep + reduce_parts + finalize per iteration
    n_enq = (iters // args.graph + (1 if iters % args.graph else 0)) if args.graph else 1
    launches_per_step = (4 + n_enq + 2 * iters) if world == 1 else (5 + 3 * iters)  # peer loop: sweep, reduce+push, finalize
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
                   "rows_per_gpu": n_loc, "kernel": kernel_desc,
                   "launch": ("one persistent launch per solve" if resident else f"CUDA graph, {args.graph} iterations per replay" if args.graph else "eager launches")
                   if world == 1 else f"loop={args.loop}", "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                   "l2": l2_note},
        "hbm_gbs": achieved * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _traffic(kernel_desc, n_loc, m), "peak_source": peak_src,
                     "note": "achieved = 4*n_local*m bytes per iteration / (step time / iterations); the step time "
                             "includes the finalize kernel and, for N>1, the exchange of the column sums (peer-memory push or NCCL)"},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "pinned host embeddings -> H2D -> cost construction -> solve -> potentials D2H"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
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
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--cpu-iters", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sampler", action="store_true", help="diagnostic: do not sample clocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
