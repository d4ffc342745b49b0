import sys, json, time
sys.path[:0] = ["/root/repo", "/root/repo/ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200"]
import torch
from b200ot import ops
dev = torch.device("cuda", 0)
B = 4096
gen = torch.Generator(device="cpu").manual_seed(20251118 + 1)
Xb = torch.randn(B, 64, 512, generator=gen); Yb = torch.randn(B, 64, 512, generator=gen) + 0.5 * torch.randn(B, 1, 512, generator=gen)
Xb = (Xb / Xb.norm(dim=2, keepdim=True)).to(dev); Yb = (Yb / Yb.norm(dim=2, keepdim=True)).to(dev)
ad = torch.full((64,), 1 / 64, device=dev)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
out = {}
out["it1_ms"] = timed(lambda: ops.sinkhorn_batched(ad, ad, 0.05, X=Xb, Y=Yb, max_iter=1, tol=0.0))
out["it201_ms"] = timed(lambda: ops.sinkhorn_batched(ad, ad, 0.05, X=Xb, Y=Yb, max_iter=201, tol=0.0))
C3 = torch.rand(B, 64, 64, device=dev) * 2
out["C3_it1_ms"] = timed(lambda: ops.sinkhorn_batched(ad, ad, 0.05, C3=C3, max_iter=1, tol=0.0))
out["C3_it201_ms"] = timed(lambda: ops.sinkhorn_batched(ad, ad, 0.05, C3=C3, max_iter=201, tol=0.0))
print(json.dumps(out))
