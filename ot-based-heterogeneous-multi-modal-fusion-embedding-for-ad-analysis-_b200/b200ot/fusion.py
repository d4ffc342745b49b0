"""OT fusion head: plan application + token attention fusion + cosine OT loss.

Mirrors the OT part of ``MultimodalMRI_PET_OT.forward`` (MRI_PET_OT_OT_per_epoch_attn.py:723-753,
per-step variant MRI_PET_OT_nojax.py:701-725):

    ot_mri_from_pet = pet_feat @ T.t()                         (:728)   -> plan-free apply_plan kernel
    tokens = [mri_feat, ot_mri_from_pet, pet2mri(pet_feat)]    (:731-738)
    attn_out = SelfAttentionBlock(tokens).mean over tokens     (:739-740) -> token-attention kernel
    ot_loss = 1 - mean cos(mri_fused, ot_mri_from_pet)         (:751, :552-560) -> cosine-loss kernel

``SelfAttentionBlock`` keeps the reference's parameter names (``self_attn.in_proj_weight`` ...,
``norm1``, ``ffn.0``, ``ffn.3``, ``norm2``), so a reference ``state_dict`` loads unchanged.  The dense
projections are library GEMMs (``F.linear``); the S = 3 softmax(QK^T)V core is the CUDA kernel
``b200ot_token_attention_{fwd,bwd}``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from ._lib import check
from .torch_ops import apply_plan


class TokenAttention(torch.autograd.Function):
    """softmax(QK^T/sqrt(dh)) V over S <= 4 tokens; qkv (S, B, 3E) -> (S, B, E)."""

    @staticmethod
    def forward(ctx, qkv, num_heads, keep_mask, keep_scale):
        lib = _lib.load()
        qkv = qkv.float().contiguous()
        S, B, E3 = qkv.shape
        E = E3 // 3
        out = torch.empty((S, B, E), dtype=torch.float32, device=qkv.device)
        probs = torch.empty((B, num_heads, S, S), dtype=torch.float32, device=qkv.device)
        check(lib.b200ot_token_attention_fwd(ops._ptr(qkv), S, B, E, num_heads, ops._ptr(keep_mask), float(keep_scale),
                                             ops._ptr(out), ops._ptr(probs), ops._stream()),
              "b200ot_token_attention_fwd")
        ctx.save_for_backward(qkv, probs, keep_mask)
        ctx.num_heads, ctx.keep_scale = num_heads, float(keep_scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        qkv, probs, keep_mask = ctx.saved_tensors
        S, B, E3 = qkv.shape
        dqkv = torch.empty_like(qkv)
        check(lib.b200ot_token_attention_bwd(ops._ptr(qkv), ops._ptr(probs), ops._ptr(keep_mask), ctx.keep_scale,
                                             ops._ptr(dout.float().contiguous()), S, B, E3 // 3, ctx.num_heads,
                                             ops._ptr(dqkv), ops._stream()), "b200ot_token_attention_bwd")
        return dqkv, None, None, None


class SelfAttentionBlock(nn.Module):
    """Transformer encoder block for feature fusion (MRI_PET_OT_OT_per_epoch_attn.py:523-549)."""

    def __init__(self, embed_dim=2048, num_heads=8, ff_dim=2048, dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=False)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.dropout1 = nn.Dropout(dropout)
        self.ffn = nn.Sequential(nn.Linear(embed_dim, ff_dim), nn.ReLU(inplace=True), nn.Dropout(dropout),
                                 nn.Linear(ff_dim, embed_dim))
        self.norm2 = nn.LayerNorm(embed_dim)
        self.dropout2 = nn.Dropout(dropout)
        self.num_heads = num_heads
        self.attn_dropout = dropout

    def _attention(self, x):
        mha = self.self_attn
        qkv = F.linear(x, mha.in_proj_weight, mha.in_proj_bias)  # (S, B, 3E)
        keep, scale = None, 1.0
        if self.training and self.attn_dropout > 0:
            S, B, _ = x.shape
            keep = (torch.rand((B, self.num_heads, S, S), device=x.device) >= self.attn_dropout).float()
            scale = 1.0 / (1.0 - self.attn_dropout)
        ctxv = TokenAttention.apply(qkv, self.num_heads, keep, scale)
        return F.linear(ctxv, mha.out_proj.weight, mha.out_proj.bias)

    def forward(self, x):
        x = self.norm1(x + self.dropout1(self._attention(x)))
        return self.norm2(x + self.dropout2(self.ffn(x)))


class CosineLoss(torch.autograd.Function):
    """1 - mean_i cos(x_i, y_i) (MRI_PET_OT_nojax.py:552-560); value from the CUDA kernel, gradient in closed form."""

    @staticmethod
    def forward(ctx, x, y):
        xd, yd = x.detach().float().contiguous(), y.detach().float().contiguous()
        ctx.save_for_backward(xd, yd)
        return ops.cosine_loss(xd, yd).reshape(()).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        nx = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        ny = y.norm(dim=1, keepdim=True).clamp_min(1e-12)
        xh, yh = x / nx, y / ny
        c = (xh * yh).sum(1, keepdim=True)
        B = x.shape[0]
        dx = -(yh - c * xh) / nx / B
        dy = -(xh - c * yh) / ny / B
        return g * dx, g * dy


def cosine_loss(x, y):
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if y.dim() == 1:
        y = y.unsqueeze(0)
    return CosineLoss.apply(x, y)


class OTFusionHead(nn.Module):
    """Steps 3b-4 and the OT loss of the reference forward.  ``plan`` is either a dense (d, d) tensor
    (the reference's ``T_feature_pet2mri``) or a ``(C, f, g, eps)`` tuple of the engine's potentials,
    in which case the plan is never materialised."""

    def __init__(self, feature_dim=512, num_heads=8, ff_dim=None, dropout=0.1):
        super().__init__()
        self.attention_mri = SelfAttentionBlock(feature_dim, num_heads, ff_dim or feature_dim, dropout)

    def forward(self, mri_feat, pet_feat, pet_to_mri, mri_fused, plan, training=False):
        if isinstance(plan, torch.Tensor):
            ot_mri_from_pet = torch.matmul(pet_feat, plan.t())
        else:
            C, f, g, eps = plan
            # (pet @ T.t())[b, k] = sum_l T[k, l] pet[b, l]  ==  (T @ pet^T)^T
            ot_mri_from_pet = apply_plan(C, f, g, eps, pet_feat.t().contiguous()).t()
        tokens = torch.stack([mri_feat, ot_mri_from_pet, pet_to_mri], dim=0)
        attn_out = self.attention_mri(tokens).transpose(0, 1).mean(dim=1)
        ot_loss = torch.zeros((), device=mri_feat.device)
        if training:
            ot_loss = cosine_loss(mri_fused, ot_mri_from_pet)
            if torch.isnan(ot_loss):
                ot_loss = torch.zeros((), device=mri_feat.device)
        return attn_out, ot_mri_from_pet, ot_loss
