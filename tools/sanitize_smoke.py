#!/usr/bin/env python
"""Small invocations of every solver kernel family in one short process (a quick "does every kernel family still
run" check, and the workload to put under compute-sanitizer where that tool is available -- it is closed on the
round-1 GPU pool, so this script has only been run plain):

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py

Shapes are tiny, but they reach the resident kernel in both ring modes, the
single-sweep cluster kernel, the robust kernels, the peer-exchange kernels (peers emulated on one GPU), the batched
and Gromov-Wasserstein one-CTA-per-problem kernels and the epilogues.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]

import numpy as np
import torch

import b200ot
from b200ot import ops, sharded

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)


def problem(n, m):
    C = torch.tensor(rng.random((n, m)), dtype=torch.float32, device=dev) * 2
    a = torch.full((n,), 1.0 / n, device=dev)
    b = torch.full((m,), 1.0 / m, device=dev)
    return ops.aligned_copy(C), a, b


for (n, m, flag) in ((9, 12, "1"), (300, 1024, "1"), (2000, 260, "1"), (300, 1024, "0"), (40, 12288, "0"), (33, 31, "1")):
    os.environ["B200OT_RESIDENT"] = flag
    C, a, b = problem(n, m)
    f, g, info = ops.sinkhorn_potentials(C, a, b, 0.1, max_iter=12, tol=0.0, check_every=5)
    assert info["n_iter"] == 12 and info["status"] == 0, info
    P = ops.plan(C, f, g, 0.1)
    assert torch.isfinite(P).all()
    ops.apply_plan(C, f, g, 0.1, torch.randn(m, 8, device=dev), normalise=True)
    ops.ot_cost(C, f, g, 0.1)
os.environ.pop("B200OT_RESIDENT")

# peer exchange, three shards on one GPU
n, m, shards = 90, 512, 3
C, a, b = problem(n, m)
prm = ops.make_params(0.1, 6, 0.0, 5, 0, "l1", False, "auto")
bufs = [torch.zeros(sharded.PeerExchange.nbytes(shards, m), dtype=torch.uint8, device=dev) for _ in range(shards)]
ks, pes = [], []
for r in range(shards):
    lo, hi = sharded.row_range(n, shards, r)
    ks.append(sharded.CudaShardKernels(C[lo:hi], a[lo:hi].contiguous(), b, prm))
    pes.append(sharded.PeerExchange(m, local_bufs=bufs, rank=r))
for k, pe in zip(ks, pes):
    k.setup()
    pe.next_epoch()
for pro in (True, False, False, False):
    for k, pe in zip(ks, pes):
        k.push(pe, pro)
    for k, pe in zip(ks, pes):
        k.finalize_peer(pe, pro)
assert ks[0].finish()[2]["n_iter"] == 3

# one-CTA-per-problem kernels
X = torch.randn(3, 20, 16, device=dev)
Y = torch.randn(3, 24, 16, device=dev)
ops.sinkhorn_batched(torch.full((20,), 0.05, device=dev), torch.full((24,), 1 / 24, device=dev), 0.5, X=X, Y=Y,
                     max_iter=20, tol=0.0)
Ts, info = ops.egw_batched([torch.randn(13, 6, device=dev), torch.randn(64, 6, device=dev)],
                           [torch.randn(9, 5, device=dev), torch.randn(64, 5, device=dev)], eps=5e-2, gw_max_iter=6)
assert all(torch.isfinite(t).all() for t in Ts)
ops.cost_matrix(torch.randn(130, 70, device=dev), torch.randn(257, 70, device=dev))
torch.cuda.synchronize()
print("sanitize smoke ok")
