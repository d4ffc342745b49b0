#!/usr/bin/env python
"""Time / profile the per-label entropic Gromov-Wasserstein kernel (one CTA per label): 3 labels of 64 x 64 samples,
d = 512, eps = 5e-3 (the per-epoch coupling of MRI_PET_OT_OT_per_epoch_attn.py:129-186)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]
import torch

from b200ot import ops

dev = torch.device("cuda", 0)
gen = torch.Generator(device="cpu").manual_seed(5)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Xs = [torch.randn(64, 512, generator=gen).abs().to(dev) for _ in range(L)]
Ys = [(x.cpu() @ torch.linalg.qr(torch.randn(512, 512, generator=gen))[0] + 0.01 * torch.randn(64, 512, generator=gen)).to(dev)
      for x in Xs]
for _ in range(2):
    Ts, info = ops.egw_batched(Xs, Ys, 5e-3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
Ts, info = ops.egw_batched(Xs, Ys, 5e-3)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"labels": L, "ms": e0.elapsed_time(e1), "outer": info["n_iters_outer"].tolist(),
                  "inner": info["inner_iterations"].tolist()}))
