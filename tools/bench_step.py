#!/usr/bin/env python
"""BASELINE config 5: OT share of an end-to-end training step (informational).

Dual 3-D ResNet-18 encoders ([2,2,2,2] basic blocks, the reference stem of MRI_PET_OT_nojax.py:466-473) on
synthetic 96^3 MRI / PET volumes, batch 32, feature_dim 512, per-step OT exactly as MRI_PET_OT_nojax.py:679-725:
features -> feature coupling (Ts = I/B, eps = 1e-2, numItermax 2000) -> row-normalised plan -> pet @ T^T ->
cosine OT loss -> backward -> AdamW.  The encoders are stock cuDNN (out of scope); what is
compared is the OT leg: (a) the reference's way -- .detach().cpu().numpy(), float64 CPU solve (oracle port),
torch.from_numpy(...).to(device) -- against (b) b200ot on device tensors.

    python tools/bench_step.py > profiles/r01_step_share.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")]

import numpy as np
import torch
import torch.nn as nn

import b200ot
from b200ot.fusion import cosine_loss
from oracle import ot_oracle as orc


class BasicBlock3D(nn.Module):
    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.c1 = nn.Conv3d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm3d(cout)
        self.c2 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm3d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv3d(cin, cout, 1, stride, bias=False), nn.BatchNorm3d(cout))

    def forward(self, x):
        y = torch.relu(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))
        return torch.relu(y + (x if self.down is None else self.down(x)))


class ResNet18_3D(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv3d(1, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3), bias=False), nn.BatchNorm3d(64),
                                  nn.ReLU(inplace=True))
        self.pool = nn.MaxPool3d(3, 2, 1)
        layers, cin = [], 64
        for cout, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
            layers += [BasicBlock3D(cin, cout, stride), BasicBlock3D(cout, cout, 1)]
            cin = cout
        self.layers = nn.Sequential(*layers)
        self.avg = nn.AdaptiveAvgPool3d(1)

    def forward(self, x):
        return torch.flatten(self.avg(self.layers(self.pool(self.stem(x)))), 1)


class StepModel(nn.Module):
    """MultimodalMRI_PET_OT (MRI_PET_OT_nojax.py:563-725) at depth 18: two backbones, the cross-modality projections,
    the fusion MLPs, the 1-token self-attention block (b200ot's SelfAttentionBlock: the reference's parameter names,
    token-attention kernel), the classifier, cross-entropy + per-step OT loss."""

    def __init__(self, d=512, num_classes=3):
        super().__init__()
        from b200ot.fusion import SelfAttentionBlock
        self.mri_backbone, self.pet_backbone = ResNet18_3D(), ResNet18_3D()

        def mlp(i, h, o):
            return nn.Sequential(nn.Linear(i, h), nn.ReLU(), nn.Dropout(0.3), nn.Linear(h, o))
        self.mri2pet, self.pet2mri = mlp(d, 2 * d, d), mlp(d, 2 * d, d)
        self.mri_fusion, self.pet_fusion = mlp(2 * d, d, d), mlp(2 * d, d, d)
        self.attention_mri = SelfAttentionBlock(embed_dim=d, num_heads=8, ff_dim=d, dropout=0.1)
        self.fc = nn.Linear(2 * d, num_classes)
        self.ce_loss = nn.CrossEntropyLoss()

    def features(self, Xm, Xp):
        mri_feat, pet_feat = self.mri_backbone(Xm), self.pet_backbone(Xp)
        mri_fused = self.mri_fusion(torch.cat([mri_feat, self.mri2pet(mri_feat)], dim=1))
        pet_fused = self.pet_fusion(torch.cat([pet_feat, self.pet2mri(pet_feat)], dim=1))
        return mri_fused, pet_fused

    def losses(self, mri_fused, pet_fused, y, T):
        attn_out = self.attention_mri(mri_fused.unsqueeze(0)).squeeze(0)
        logits = self.fc(torch.cat([attn_out, pet_fused], dim=1))
        ce = self.ce_loss(logits, y)
        ot_mri_from_pet = torch.matmul(pet_fused, T.t())          # :718
        ot = cosine_loss(mri_fused, ot_mri_from_pet)              # :721 (CUDA kernel + closed-form gradient)
        if torch.isnan(ot):
            ot = torch.zeros((), device=ce.device)
        return ce + ot


def guard_rownorm(T):
    T = torch.nan_to_num(T, nan=1e-8)
    rs = T.sum(dim=1, keepdim=True)
    return T / torch.where(rs == 0, torch.full_like(rs, 1e-8), rs)


def measure(B=32, S=96, steps=5, warm=2, dev=None):
    """OT share of the end-to-end training step, reference CPU path vs device path (same model, same inputs)."""
    dev = dev or torch.device("cuda", 0)
    torch.manual_seed(0)
    model = StepModel().to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    Xm = torch.randn(B, 1, S, S, S, device=dev)
    Xp = torch.randn(B, 1, S, S, S, device=dev)
    y = torch.randint(0, 3, (B,), device=dev)

    def ot_cpu(mri, pet):
        # the reference's way (:683-715): device -> host, float64 CPU solve, host -> device, guard in torch
        mri_np, pet_np = mri.detach().cpu().numpy(), pet.detach().cpu().numpy()
        Tv, _ = orc.get_feature_coupling_pot(({0: mri_np}, {0: pet_np}), {0: np.eye(B) / B}, eps=1e-2)
        return guard_rownorm(torch.from_numpy(Tv).float().to(dev))

    def ot_gpu(mri, pet):
        # the whole block on device tensors, guard fused into the kernel that writes the plan
        return b200ot.per_step_feature_plan(mri, pet, eps=1e-2)

    def step(ot_fn, timing):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        mri, pet = model.features(Xm, Xp)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        T = ot_fn(mri, pet)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        loss = model.losses(mri, pet, y, T)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        timing.append((t1 - t0, t2 - t1, t3 - t2))

    out = {"config": f"MultimodalMRI_PET_OT at depth 18 (dual 3-D ResNet-18 + projections + fusion MLPs + attention block "
                     f"+ classifier), batch {B}, {S}^3 volumes, feature_dim 512, per-step feature OT 512x512 "
                     f"(Ts = I/B, eps = 1e-2, numItermax 2000, POT rule), CE + OT loss, AdamW "
                     f"(BASELINE configs[4]; encoders are stock cuDNN, out of scope)"}
    for name, fn in (("reference_cpu_path", ot_cpu), ("b200ot_device_path", ot_gpu)):
        tm = []
        for _ in range(warm + steps):
            step(fn, tm)
        tm = np.array(tm[warm:])
        enc, ot, rest = tm.mean(0)
        out[name] = {"forward_to_features_ms": 1e3 * enc, "ot_ms": 1e3 * ot, "loss_bwd_opt_ms": 1e3 * rest,
                     "step_ms": 1e3 * (enc + ot + rest), "ot_share": float(ot / (enc + ot + rest))}
    out["ot_speedup"] = out["reference_cpu_path"]["ot_ms"] / out["b200ot_device_path"]["ot_ms"]
    return out


def main():
    B, S = int(os.environ.get("STEP_BATCH", 32)), int(os.environ.get("STEP_SIDE", 96))
    print(json.dumps(measure(B, S), indent=1))


if __name__ == "__main__":
    main()
