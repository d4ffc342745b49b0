"""GPU tests of the autograd surface: envelope gradient of the OT loss, plan application with gradient
to V, the token-attention kernel against PyTorch's own nn.MultiheadAttention (fp32 reference of the same
op), and the fusion head against the reference forward's arithmetic restated with stock torch modules."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import ot_oracle as orc

pytestmark = pytest.mark.gpu


def _unrolled_dual_value(x, y, a, b, eps, iters):
    """float64 torch: log-domain Sinkhorn unrolled, returns <a,f> + <b,g> (autograd flows through everything)."""
    C = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2 * x @ y.T
    f = torch.zeros_like(a)
    g = torch.zeros_like(b)
    for _ in range(iters):
        g = eps * torch.log(b) - eps * torch.logsumexp((f[:, None] - C) / eps, dim=0)
        f = eps * torch.log(a) - eps * torch.logsumexp((g[None, :] - C) / eps, dim=1)
    return (a * f).sum() + (b * g).sum(), C, f, g


def test_ot_loss_envelope_gradient_matches_unrolled_autograd(cuda_dev):
    from b200ot.torch_ops import ot_loss
    n, m, d, eps = 96, 128, 32, 0.1
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=6)
    a = np.ones(n) / n
    b = np.ones(m) / m
    xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    yt = torch.tensor(Y, dtype=torch.float64, requires_grad=True)
    val, C, f, g = _unrolled_dual_value(xt, yt, torch.tensor(a), torch.tensor(b), eps, 400)
    val.backward()
    xd = torch.tensor(X, device=cuda_dev, requires_grad=True)
    yd = torch.tensor(Y, device=cuda_dev, requires_grad=True)
    loss = ot_loss(xd, yd, eps=eps, max_iter=400, tol=0.0, value="dual")
    loss.backward()
    assert abs(loss.item() - val.item()) < 1e-4 * abs(val.item())
    np.testing.assert_allclose(xd.grad.cpu().numpy(), xt.grad.numpy(), rtol=0, atol=1e-4 * np.abs(xt.grad.numpy()).max())
    np.testing.assert_allclose(yd.grad.cpu().numpy(), yt.grad.numpy(), rtol=0, atol=1e-4 * np.abs(yt.grad.numpy()).max())
    # closed form of the oracle (envelope_grads) agrees as well
    P = orc.plan_from_potentials(C.detach().numpy(), f.detach().numpy(), g.detach().numpy(), eps)
    dX, dY = orc.envelope_grads(X, Y, P)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), dX, rtol=0, atol=1e-4 * np.abs(dX).max())
    # primal value <P, C>
    lp = ot_loss(xd.detach(), yd.detach(), eps=eps, max_iter=400, tol=0.0, value="primal")
    assert abs(lp.item() - orc.ot_cost(P, C.detach().numpy())) < 1e-4 * orc.ot_cost(P, C.detach().numpy())


def test_apply_plan_gradient_flows_to_values(cuda_dev):
    from b200ot import ops
    from b200ot.torch_ops import apply_plan
    rng = np.random.default_rng(2)
    n, m, dv, eps = 70, 90, 24, 0.2
    C = rng.random((n, m))
    a = np.ones(n) / n
    b = np.ones(m) / m
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=50, tol=0.0, log=True)
    Cd = torch.tensor(C, dtype=torch.float32, device=cuda_dev)
    fd = torch.tensor(lg["f"], dtype=torch.float32, device=cuda_dev)
    gd = torch.tensor(lg["g"], dtype=torch.float32, device=cuda_dev)
    V = rng.standard_normal((m, dv))
    W = rng.standard_normal((n, dv))
    for normalise in (False, True):
        Vd = torch.tensor(V, dtype=torch.float32, device=cuda_dev, requires_grad=True)
        Z = apply_plan(Cd, fd, gd, eps, Vd, normalise=normalise)
        (Z * torch.tensor(W, dtype=torch.float32, device=cuda_dev)).sum().backward()
        Pt = torch.tensor(Pref)
        if normalise:
            Pt = Pt / Pt.sum(1, keepdim=True)
        Vt = torch.tensor(V, requires_grad=True)
        ((Pt @ Vt) * torch.tensor(W)).sum().backward()
        np.testing.assert_allclose(Z.detach().cpu().numpy(), (Pt @ Vt).detach().numpy(), rtol=0,
                                   atol=1e-4 * float((Pt @ Vt).abs().max()))
        np.testing.assert_allclose(Vd.grad.cpu().numpy(), Vt.grad.numpy(), rtol=0, atol=1e-4 * float(Vt.grad.abs().max()))


@pytest.mark.parametrize("S,B,E,H", [(3, 4, 512, 8), (1, 5, 512, 8), (3, 32, 2048, 8), (4, 3, 96, 3)])
def test_token_attention_matches_torch_mha(cuda_dev, S, B, E, H):
    from b200ot.fusion import SelfAttentionBlock
    torch.manual_seed(S * 100 + B)
    blk = SelfAttentionBlock(E, H, ff_dim=E, dropout=0.1).to(cuda_dev).eval()
    x = torch.randn(S, B, E, device=cuda_dev, requires_grad=True)
    x_ref = x.detach().clone().requires_grad_(True)
    # stock PyTorch restatement of the reference block (MRI_PET_OT_OT_per_epoch_attn.py:523-549)
    def ref_forward(t):
        attn, _ = blk.self_attn(t, t, t)
        t = blk.norm1(t + attn)
        return blk.norm2(t + blk.ffn(t))
    out = blk(x)
    out_ref = ref_forward(x_ref)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    gx = x.grad.clone()
    gparams = [p.grad.clone() for p in blk.parameters()]
    blk.zero_grad()
    (out_ref * w).sum().backward()
    assert torch.allclose(out, out_ref, rtol=1e-4, atol=2e-5)
    assert torch.allclose(gx, x_ref.grad, rtol=1e-3, atol=2e-5 * float(x_ref.grad.abs().max() + 1))
    for g1, p in zip(gparams, blk.parameters()):
        assert torch.allclose(g1, p.grad, rtol=1e-3, atol=2e-5 * float(p.grad.abs().max() + 1))


def test_token_attention_dropout_mask(cuda_dev):
    from b200ot.fusion import TokenAttention
    torch.manual_seed(0)
    S, B, E, H = 3, 6, 64, 4
    qkv = torch.randn(S, B, 3 * E, device=cuda_dev, requires_grad=True)
    keep = (torch.rand(B, H, S, S, device=cuda_dev) > 0.3).float()
    out = TokenAttention.apply(qkv, H, keep, 1 / 0.7)
    q, k, v = qkv.detach().clone().requires_grad_(True).chunk(3, dim=-1)
    qr = qkv.detach().clone().requires_grad_(True)
    q, k, v = qr.chunk(3, dim=-1)
    def heads(t):
        return t.reshape(S, B, H, E // H).permute(1, 2, 0, 3)
    p = torch.softmax(heads(q) @ heads(k).transpose(-1, -2) / (E // H) ** 0.5, dim=-1) * keep / 0.7
    ref = (p @ heads(v)).permute(2, 0, 1, 3).reshape(S, B, E)
    w = torch.randn_like(ref)
    (out * w).sum().backward()
    (ref * w).sum().backward()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(qkv.grad, qr.grad, rtol=1e-3, atol=1e-5)


def test_fusion_head_matches_reference_forward_arithmetic(cuda_dev):
    """Dense-plan and plan-free (potentials) forms of the head agree with the reference's arithmetic."""
    from b200ot import ops
    from b200ot.fusion import OTFusionHead
    torch.manual_seed(3)
    d, B, eps = 512, 8, 0.05
    head = OTFusionHead(d, 8, dropout=0.1).to(cuda_dev).eval()
    mri_feat = torch.randn(B, d, device=cuda_dev)
    pet_feat = torch.randn(B, d, device=cuda_dev, requires_grad=True)
    pet_to_mri = torch.randn(B, d, device=cuda_dev)
    mri_fused = torch.randn(B, d, device=cuda_dev)
    # a feature-feature plan from the engine itself
    Xs, Ys = orc.synthetic_embeddings(64, 64, d, config_index=0)
    xs, ys = torch.tensor(Xs, device=cuda_dev), torch.tensor(Ys, device=cuda_dev)
    Ts = torch.eye(64, device=cuda_dev) / 64
    M = ops.fot_cost(xs, ys, Ts, Ts.sum(1), Ts.sum(0))
    a = torch.full((d,), 1.0 / d, device=cuda_dev)
    f, g, _ = ops.sinkhorn_potentials(M, a, a, 1e-2, max_iter=50, tol=0.0)
    T = ops.plan(M, f, g, 1e-2)
    attn1, z1, l1 = head(mri_feat, pet_feat, pet_to_mri, mri_fused, T, training=True)
    attn2, z2, l2 = head(mri_feat, pet_feat, pet_to_mri, mri_fused, (M, f, g, 1e-2), training=True)
    assert torch.allclose(z1, z2, rtol=1e-4, atol=1e-6 * float(z1.abs().max()) + 1e-9)
    assert torch.allclose(attn1, attn2, rtol=1e-4, atol=1e-5)
    assert abs(l1.item() - l2.item()) < 1e-5
    # reference arithmetic: pet @ T.t(), cosine loss with F.normalize / F.cosine_similarity
    z_ref = pet_feat @ T.t()
    l_ref = 1 - F.cosine_similarity(F.normalize(mri_fused, p=2, dim=1), F.normalize(z_ref, p=2, dim=1)).mean()
    assert abs(l1.item() - l_ref.item()) < 1e-5
    # gradient reaches pet_feat through the plan application (the only gradient path of the reference)
    (attn2.sum() + l2).backward()
    g2 = pet_feat.grad.clone()
    pet_feat.grad = None
    (attn1.sum() + l1).backward()
    assert torch.allclose(g2, pet_feat.grad, rtol=1e-3, atol=1e-5 * float(pet_feat.grad.abs().max()))


def _unrolled_primal_value(x, y, a, b, eps, iters):
    """float64 torch: log-domain Sinkhorn unrolled, returns <P, C> (autograd flows through the iterations)."""
    C = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2 * x @ y.T
    f = torch.zeros_like(a)
    g = torch.zeros_like(b)
    for _ in range(iters):
        g = eps * torch.log(b) - eps * torch.logsumexp((f[:, None] - C) / eps, dim=0)
        f = eps * torch.log(a) - eps * torch.logsumexp((g[None, :] - C) / eps, dim=1)
    P = torch.exp((f[:, None] + g[None, :] - C) / eps)
    return (P * C).sum()


def test_implicit_gradient_of_the_transport_cost_matches_unrolled_autograd(cuda_dev):
    """ot_loss(value="primal", grad="implicit"): d<P, C>/dx through the Sinkhorn fixed point (implicit function
    theorem; CG with plan-free products + one weighted launch of the tcgen05 kernel) against float64 autograd
    through 600 unrolled iterations.  The fixed-plan (envelope) gradient is visibly different for this value."""
    from b200ot.torch_ops import ot_loss
    n, m, d, eps = 160, 192, 24, 0.2
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=9)
    a = torch.full((n,), 1.0 / n, dtype=torch.float64)
    b = torch.full((m,), 1.0 / m, dtype=torch.float64)
    xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    yt = torch.tensor(Y, dtype=torch.float64, requires_grad=True)
    val = _unrolled_primal_value(xt, yt, a, b, eps, 600)
    val.backward()
    xd = torch.tensor(X, device=cuda_dev, requires_grad=True)
    yd = torch.tensor(Y, device=cuda_dev, requires_grad=True)
    loss = ot_loss(xd, yd, eps=eps, max_iter=600, tol=0.0, value="primal", grad="implicit")
    loss.backward()
    assert abs(loss.item() - val.item()) < 1e-4 * abs(val.item())
    gx, gy = xt.grad.numpy(), yt.grad.numpy()
    np.testing.assert_allclose(xd.grad.cpu().numpy(), gx, rtol=0, atol=3e-3 * np.abs(gx).max())
    np.testing.assert_allclose(yd.grad.cpu().numpy(), gy, rtol=0, atol=3e-3 * np.abs(gy).max())
    # the envelope gradient of the same value ignores dP/dx: it must NOT agree to that tolerance
    xe = torch.tensor(X, device=cuda_dev, requires_grad=True)
    ot_loss(xe, yd.detach(), eps=eps, max_iter=600, tol=0.0, value="primal").backward()
    assert np.abs(xe.grad.cpu().numpy() - gx).max() > 3e-2 * np.abs(gx).max()


def test_weighted_plan_products_match_float64(cuda_dev):
    """b200ot_apply_plan_tc_weighted: W = P o (w0 + w1 C + wrow (+) wcol), both forms, against NumPy."""
    from b200ot import ops
    rng = np.random.default_rng(31)
    n, m, dv = 300, 520, 40
    C = rng.random((n, m))
    f = rng.standard_normal(n) * 0.05
    g = rng.standard_normal(m) * 0.05
    eps = 0.3
    P = orc.plan_from_potentials(C, f, g, eps)
    wr, wc = rng.standard_normal(n), rng.standard_normal(m)
    Wm = P * (0.7 - 1.3 * C + wr[:, None] + wc[None, :])
    V, U = rng.standard_normal((m, dv)), rng.standard_normal((n, dv))
    dev = cuda_dev
    t = lambda z: torch.tensor(z, dtype=torch.float32, device=dev)  # noqa: E731
    Z, rs = ops.apply_plan(t(C), t(f), t(g), eps, t(V), weights=(0.7, -1.3, t(wr), t(wc)), return_rowsum=True)
    np.testing.assert_allclose(Z.double().cpu().numpy(), Wm @ V, rtol=0, atol=1e-4 * np.abs(Wm @ V).max())
    np.testing.assert_allclose(rs.double().cpu().numpy(), Wm.sum(1), rtol=0, atol=1e-4 * np.abs(Wm.sum(1)).max())
    Zt, cs = ops.apply_plan(t(C), t(f), t(g), eps, t(U), transpose=True, weights=(0.7, -1.3, t(wr), t(wc)),
                            return_rowsum=True)
    np.testing.assert_allclose(Zt.double().cpu().numpy(), Wm.T @ U, rtol=0, atol=1e-4 * np.abs(Wm.T @ U).max())
    np.testing.assert_allclose(cs.double().cpu().numpy(), Wm.sum(0), rtol=0, atol=1e-4 * np.abs(Wm.sum(0)).max())
