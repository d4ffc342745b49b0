#!/usr/bin/env python
"""One GPU, the shape of one rank of the 8-GPU benchmark (8192 rows x 65536 columns): per-iteration time of the
forms of the row-sharded peer loop with a single-rank exchange (push and poll hit the same buffer, so what is
measured is launch structure and fold cost, not NVLink).  CUDA events, 200 iterations each."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch  # noqa: E402

from b200ot import ops, sharded  # noqa: E402

dev = torch.device("cuda", 0)
n_local, m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 65536
iters = 200
gen = torch.Generator(device="cpu").manual_seed(3)
X = torch.randn(n_local, 64, generator=gen)
Y = torch.randn(m, 64, generator=gen)
X = (X / X.norm(dim=1, keepdim=True)).to(dev)
Y = (Y / Y.norm(dim=1, keepdim=True)).to(dev)
C = ops.cost_matrix(X, Y)
a = torch.full((n_local,), 1.0 / (8 * n_local), device=dev)
b = torch.full((m,), 1.0 / m, device=dev)
prm = ops.make_params(0.05, 10 ** 6, 0.0, 10, 1, "l2", False, "auto")
buf = torch.zeros(sharded.PeerExchange.nbytes(1, m), dtype=torch.uint8, device=dev)


def timed(fn):
    fn(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(iters)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


out = {"n_local": n_local, "m": m, "bytes_per_iteration": 4.0 * n_local * m}
epoch = 1
for name, env in (("sweep_plus_tail_default", {}), ("three_launches_PEER_TAIL_0", {"B200OT_PEER_TAIL": "0"}),
                  ("persistent_fused_FUSE_1", {"B200OT_FUSE": "1"})):
    os.environ.update(env)
    os.environ["B200OT_RESIDENT"] = "0"
    k = sharded.CudaShardKernels(C, a, b, prm)
    pe = sharded.PeerExchange(m, local_bufs=[buf], rank=0)
    epoch += 1
    pe.epoch = epoch
    k.setup()
    k.push(pe, True)
    k.finalize_peer(pe, True)
    out[name + "_us"] = timed(lambda it: k.run_peer(it, pe))
    assert k.flags()["bad"] == 0
    for key in env:
        del os.environ[key]
k = sharded.CudaShardKernels(C, a, b, prm)
k.setup()
k.finalize(k.prologue(), True)
out["sweep_plus_reduce_parts_us"] = timed(lambda it: [k.sweep() for _ in range(it)])
out["gbs_default"] = out["bytes_per_iteration"] / out["sweep_plus_tail_default_us"] / 1e3
print(json.dumps(out))
