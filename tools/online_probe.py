#!/usr/bin/env python
"""A few iterations of the online (cost-free) solver at n = m = 65536, d = 512, for the ncu launch list."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch

from b200ot.online import OnlineSinkhorn

dev = torch.device("cuda", 0)
n = m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
its = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pb = int(sys.argv[3]) if len(sys.argv) > 3 else (2 << 30)
terms = sys.argv[4] if len(sys.argv) > 4 else None   # "f16", "f16x4" or 6 / 3 / 1 (default: the package default)
terms = int(terms) if terms is not None and terms.isdigit() else terms
gen = torch.Generator(device="cpu").manual_seed(20251121)
X = torch.randn(n, 512, generator=gen)
Y = torch.randn(m, 512, generator=gen) + 0.5 * torch.randn(1, 512, generator=gen)
X = (X / X.norm(dim=1, keepdim=True)).to(dev)
Y = (Y / Y.norm(dim=1, keepdim=True)).to(dev)
a = torch.full((n,), 1.0 / n, device=dev)
b = torch.full((m,), 1.0 / m, device=dev)
sol = OnlineSinkhorn(X, Y, a, b, 0.05, max_iter=10 ** 6, tol=0.0, panel_bytes=pb, terms=terms)
sol.start()
sol.run(1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sol.run(its)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / its
print(json.dumps({"n": n, "panel_rows": sol.panel_rows, "terms": str(terms), "products": sol.products,
                  "ms_per_iteration": ms,
                  "tensor_tflops_executed": sol.tensor_flops_per_iteration / ms / 1e9}))
