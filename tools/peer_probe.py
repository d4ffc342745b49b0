#!/usr/bin/env python
"""torchrun probe (N >= 2 GPUs of one box): the peer-memory sharded loop against the NCCL loop.

Checks that both loops stop on the same iteration with the same potentials (the fold order differs, so fp32
rounding does), that g is replicated bit for bit across ranks in the peer loop, and times both.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_probe.py 16384
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch
import torch.distributed as dist

from b200ot import ops, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lo, hi = sharded.row_range(n, world, rank)
gen = torch.Generator(device="cpu").manual_seed(1234)
X = torch.randn(n, 64, generator=gen)
Y = torch.randn(m, 64, generator=gen) + 0.5 * torch.randn(1, 64, generator=gen)
X = X / X.norm(dim=1, keepdim=True)
Y = Y / Y.norm(dim=1, keepdim=True)
C = ops.cost_matrix(X[lo:hi].to(dev), Y.to(dev))
a = torch.full((hi - lo,), 1.0 / n, device=dev)
b = torch.full((m,), 1.0 / m, device=dev)
out = {"n": n, "world": world}

peer = sharded.PeerExchange(m)
comm = sharded.NcclComm()

# ---- parity: convergence run with the ott rule
res = {}
for name, kw in (("peer", {"peer": peer}), ("nccl", {"comm": comm})):
    f, g, info = sharded.solve_sharded(C, a, b, 0.05, max_iter=500, tol=1e-3, check_every=10, check_phase=0,
                                       err_norm="l1", **kw)
    res[name] = (f, g, info)
fp, gp, ip = res["peer"]
fn, gn, inn = res["nccl"]
assert ip["n_iter"] == inn["n_iter"] and ip["converged"] and inn["converged"], (ip, inn)
assert ip["status"] == 0
torch.testing.assert_close(gp, gn, rtol=0, atol=2e-5)
torch.testing.assert_close(fp, fn, rtol=0, atol=2e-5)
gs = [torch.empty_like(gp) for _ in range(world)]
dist.all_gather(gs, gp)
assert all(torch.equal(gs[0], t) for t in gs), "g differs across ranks in the peer loop"
# a second solve on the same buffers (new epoch) gives the same bits
f2, g2, i2 = sharded.solve_sharded(C, a, b, 0.05, max_iter=500, tol=1e-3, check_every=10, check_phase=0,
                                   err_norm="l1", peer=peer)
assert torch.equal(g2, gp) and torch.equal(f2, fp) and i2["n_iter"] == ip["n_iter"]
out["parity"] = {"n_iter": ip["n_iter"], "max_abs_dg_vs_nccl": float((gp - gn).abs().max())}

# ---- timing: fixed iteration count
prm = ops.make_params(0.05, 10 ** 6, 0.0, 10, 1, "l2", False, "auto")
for name, kw in (("peer", {"peer": peer}), ("nccl", {"comm": comm})):
    k = sharded.CudaShardKernels(C, a, b, prm)
    drv = sharded.ShardedSinkhorn(k, **kw)
    drv.start()
    drv.run(20)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    drv.run(iters)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name + "_ms_per_iteration"] = float(t.item())
    assert k.flags()["bad"] == 0
# ---- where an iteration's time goes: the local sweep alone (sweep + fold of the cluster partials, no exchange,
# the state does not advance) against the full loop above
k = sharded.CudaShardKernels(C, a, b, prm)
k.setup()
k.finalize(k.prologue(), True)
for _ in range(5):
    k.sweep()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    k.sweep()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
tmax = t.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone()
dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
out["sweep_plus_fold_ms_max_over_ranks"] = float(tmax.item())
out["sweep_plus_fold_ms_min_over_ranks"] = float(tmin.item())
out["fuse"] = os.environ.get("B200OT_FUSE", "1")
out["hbm_bytes_per_iteration_per_rank"] = 4.0 * (hi - lo) * m
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
