#!/usr/bin/env python
"""torchrun probe: where does the per-iteration time of the row-sharded loop go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch, torch.distributed as dist
from b200ot import ops, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
lo, hi = sharded.row_range(n, world, rank)
C = torch.rand((hi - lo, m), device=dev) * 4
a = torch.full((hi - lo,), 1.0 / n, device=dev)
b = torch.full((m,), 1.0 / m, device=dev)
prm = ops.make_params(0.05, 100000, 0.0, 10, 1, "l2", False, "auto")
k = sharded.CudaShardKernels(C, a, b, prm)
drv = sharded.ShardedSinkhorn(k)

def timed(fn, iters, label):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(iters)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / iters
    if rank == 0:
        print(f"{label:40s} {dt*1e3:8.3f} ms/iter", flush=True)

drv.start()
timed(lambda it: drv.run(it), 20, "warmup")
timed(lambda it: drv.run(it), 100, "run: no host sync")
def windowed(w):
    def f(it):
        evs = []
        for i in range(it):
            drv.run(1)
            if i % w == w - 1:
                e = torch.cuda.Event(); e.record(); evs.append(e)
                if len(evs) > 1:
                    evs.pop(0).synchronize()
    return f
timed(windowed(8), 100, "run: event window 8")
timed(windowed(32), 100, "run: event window 32")
def sweep_only(it):
    for _ in range(it):
        k.sweep()
timed(sweep_only, 100, "sweep + reduce_parts only")
def sweep_fin(it):
    for _ in range(it):
        k.finalize(k.sweep(), False)
timed(sweep_fin, 100, "sweep + finalize (no all-reduce)")
s = torch.zeros(m, device=dev)
def ar_only(it):
    for _ in range(it):
        dist.all_reduce(s)
timed(ar_only, 200, "all_reduce of m floats only")
dist.destroy_process_group()
