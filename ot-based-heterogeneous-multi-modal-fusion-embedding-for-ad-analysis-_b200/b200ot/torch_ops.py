"""``torch.library`` custom ops and autograd functions over the C ABI.

The ops (``torch.ops.b200ot.*``) make the engine visible to PyTorch's dispatcher (fake
tensors / graph capture see shapes without running CUDA); the autograd functions put the OT
solve inside a training graph:

  * ``OTLoss``       embeddings -> entropic OT value, gradient by the envelope theorem
                     (dL/dx_i = sum_j P_ij dC_ij/dx_i with P held at the optimum).  The
                     reference never differentiates through OT (features are detached at
                     MRI_PET_OT_nojax.py:683-684); this is the north star's new capability.
  * ``ApplyPlan``    Z = [diag(1/P1)] P V with the plan constant and the gradient flowing
                     to V only -- exactly how the reference's ``pet_feat @ T.t()`` re-enters
                     autograd (MRI_PET_OT_OT_per_epoch_attn.py:728).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib, ops

_NORM_NAMES = {v: k for k, v in _lib.NORMS.items()}
_PATH_NAMES = {v: k for k, v in _lib.PATHS.items()}


@torch.library.custom_op("b200ot::cost", mutates_args=())
def cost_op(x: Tensor, y: Tensor, kind: int) -> Tensor:
    return ops.cost_matrix(x, y, kind="cosine" if kind == _lib.COST_COSINE else "sqeuclidean").contiguous()


@cost_op.register_fake
def _(x, y, kind):
    return x.new_empty((x.shape[0], y.shape[0]))


@torch.library.custom_op("b200ot::sinkhorn_fwd", mutates_args=())
def sinkhorn_fwd(C: Tensor, a: Tensor, b: Tensor, eps: float, max_iter: int, tol: float, check_every: int,
                 check_phase: int, err_norm: int, stop_inclusive: bool, path: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(f, g, stats) with stats = [n_iter, converged, err, status] as fp32."""
    f, g, info = ops.sinkhorn_potentials(ops.aligned_copy(C), a, b, eps, max_iter=max_iter, tol=tol,
                                         check_every=check_every, check_phase=check_phase,
                                         err_norm=_NORM_NAMES[err_norm], stop_inclusive=stop_inclusive,
                                         path=_PATH_NAMES[path])
    stats = torch.tensor([info["n_iter"], float(info["converged"]), info["err"], info["status"]],
                         dtype=torch.float32, device=C.device)
    return f, g, stats


@sinkhorn_fwd.register_fake
def _(C, a, b, eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive, path):
    return C.new_empty((C.shape[0],)), C.new_empty((C.shape[1],)), C.new_empty((4,))


@torch.library.custom_op("b200ot::plan", mutates_args=())
def plan_op(C: Tensor, f: Tensor, g: Tensor, eps: float) -> Tensor:
    return ops.plan(C, f, g, eps)


@plan_op.register_fake
def _(C, f, g, eps):
    return C.new_empty(C.shape)


@torch.library.custom_op("b200ot::apply_plan", mutates_args=())
def apply_plan_op(C: Tensor, f: Tensor, g: Tensor, eps: float, V: Tensor, normalise: bool, transpose: bool) -> Tensor:
    return ops.apply_plan(C, f, g, eps, V, normalise=normalise, transpose=transpose)


@apply_plan_op.register_fake
def _(C, f, g, eps, V, normalise, transpose):
    return V.new_empty((C.shape[1] if transpose else C.shape[0], V.shape[1]))


@torch.library.custom_op("b200ot::sinkhorn_bwd_envelope", mutates_args=())
def sinkhorn_bwd_envelope(C: Tensor, f: Tensor, g: Tensor, eps: float, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
    """Envelope gradients of <P, C(x, y)> for the squared-Euclidean cost at fixed P:
    dx = 2 (diag(P1) x - P y), dy = 2 (diag(P^T 1) y - P^T x): one launch of the tcgen05 plan-application
    kernel computes both products, the row / column sums of P and the combination (b200ot_envelope_bwd)."""
    return ops.envelope_bwd(C, f, g, eps, x.contiguous(), y.contiguous(), scale=2.0)


@sinkhorn_bwd_envelope.register_fake
def _(C, f, g, eps, x, y):
    return torch.empty_like(x), torch.empty_like(y)


def implicit_weights(C: Tensor, f: Tensor, g: Tensor, eps: float, tol: float = 1e-6, max_cg: int = 200):
    """Adjoint of the Sinkhorn fixed point for the transport cost L = <P, C> (implicit function theorem).

    With P = exp((f (+) g - C)/eps), r = (P o C) 1, c = (P o C)^T 1 and H = [[diag(P1), P], [P^T, diag(P^T 1)]]
    (positive semi-definite, null space (1, -1)), H (lambda, mu) = (r, c) gives
    dL/dC = P o (1 - C/eps + (lambda (+) mu)/eps).  The system is solved on its Schur complement
    S mu = c - P^T (r / P1), S = diag(P^T 1) - P^T diag(1/P1) P, by conjugate gradients; every product with P or P^T
    is one plan-free pass over C (b200ot_apply_plan / _t), the C-weighted marginals come from the weighted
    tensor-core kernel.  Returns (w0, w1, wrow, wcol) for ops.envelope_bwd(weights=...) and the CG iteration count.
    Exact only at a converged plan (P1 = a, P^T 1 = b); fp32 products bound the accuracy of the gradient to ~1e-3."""
    n, m = C.shape
    ones_m = torch.ones((m, 1), dtype=torch.float32, device=C.device)
    ones_n = torch.ones((n, 1), dtype=torch.float32, device=C.device)
    _, r = ops.apply_plan(C, f, g, eps, ones_m, return_rowsum=True, weights=(0.0, 1.0, None, None))
    _, c = ops.apply_plan(C, f, g, eps, ones_n, transpose=True, return_rowsum=True, weights=(0.0, 1.0, None, None))
    pa = ops.apply_plan(C, f, g, eps, ones_m, impl="simt").reshape(-1)                  # P 1
    pb = ops.apply_plan(C, f, g, eps, ones_n, transpose=True, impl="simt").reshape(-1)  # P^T 1
    pa_safe = torch.where(pa > 0, pa, torch.ones_like(pa))

    def Pv(v):
        return ops.apply_plan(C, f, g, eps, v.reshape(-1, 1).contiguous(), impl="simt").reshape(-1)

    def Ptv(u):
        return ops.apply_plan(C, f, g, eps, u.reshape(-1, 1).contiguous(), transpose=True, impl="simt").reshape(-1)

    def S(v):
        return pb * v - Ptv(Pv(v) / pa_safe)

    rhs = (c - Ptv(r / pa_safe)).double()
    rhs = rhs - rhs.mean()          # consistent right-hand side: orthogonal to the null space (constants)
    mu = torch.zeros(m, dtype=torch.float64, device=C.device)
    res = rhs.clone()
    d = res.clone()
    rr = float(res @ res)
    rr0 = rr
    it = 0
    while it < max_cg and rr > (tol * tol) * rr0 and rr0 > 0:
        Sd = S(d.float()).double()
        Sd = Sd - Sd.mean()
        alpha = rr / float(d @ Sd)
        mu = mu + alpha * d
        res = res - alpha * Sd
        rr_new = float(res @ res)
        d = res + (rr_new / rr) * d
        rr = rr_new
        it += 1
    mu = mu.float()
    lam = (r - Pv(mu)) / pa_safe
    return (1.0, -1.0 / eps, lam / eps, mu / eps), it


class OTLoss(torch.autograd.Function):
    """Entropic OT value between two embedding clouds (squared-Euclidean cost).

    forward:  C = cost(x, y);  (f, g) = Sinkhorn(C, a, b, eps);
              value = <P, C>                      (``value="primal"``, fot.py:137's cost)
                    | <a, f> + <b, g>             (``value="dual"``, the regularised OT value up to a constant)
    backward: envelope theorem -- P is held fixed, d value / dx = 2 (diag(P1) x - P y) and the same for y.
    """

    @staticmethod
    def forward(ctx, x, y, a, b, eps, max_iter, tol, value, grad="envelope"):
        ctx.implicit = (grad == "implicit")
        xd, yd = x.detach().float().contiguous(), y.detach().float().contiguous()
        C = ops.cost_matrix(xd, yd)
        f, g, info = ops.sinkhorn_potentials(C, a, b, eps, max_iter=max_iter, tol=tol)
        ctx.save_for_backward(C, f, g, xd, yd)
        ctx.eps = eps
        ctx.info = info
        ctx.x_dtype, ctx.y_dtype = x.dtype, y.dtype
        if value == "dual":
            out = (a * f).sum() + (b * g).sum()
        else:
            out = ops.ot_cost(C, f, g, eps).to(torch.float32).reshape(())
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        C, f, g, xd, yd = ctx.saved_tensors
        if ctx.implicit:
            w, ctx.cg_iterations = implicit_weights(C, f, g, ctx.eps)
            dx, dy = ops.envelope_bwd(C, f, g, ctx.eps, xd, yd, scale=2.0, weights=w)
        else:
            dx, dy = torch.ops.b200ot.sinkhorn_bwd_envelope(C, f, g, ctx.eps, xd, yd)
        return ((grad_out * dx).to(ctx.x_dtype), (grad_out * dy).to(ctx.y_dtype), None, None, None, None, None,
                None, None)


def ot_loss(x: Tensor, y: Tensor, a: Tensor = None, b: Tensor = None, eps: float = 0.05, max_iter: int = 1000,
            tol: float = 1e-6, value: str = "dual", grad: str = "envelope") -> Tensor:
    """Entropic OT value between two embedding clouds with its envelope gradient.

    ``value="dual"`` (default): ``<a, f> + <b, g>``, the regularised OT value up to a constant -- the quantity
    whose exact gradient the envelope theorem gives (P held at the optimum).  ``value="primal"``: the transport
    cost ``<P, C>`` (perturbot/perturbot/match/fot.py:137); its backward is the same fixed-plan gradient, i.e. a
    stop-gradient approximation that ignores dP/dx (the entropy term's contribution).  Gradients come back in
    the dtype of the inputs.  ``grad="implicit"`` (with ``value="primal"``): the gradient of ``<P, C>`` THROUGH
    the Sinkhorn fixed point by the implicit function theorem (``implicit_weights``: a conjugate-gradient solve
    whose matrix-vector products are plan-free passes over C, then one weighted launch of the tcgen05
    plan-application kernel for both embeddings); needs a converged solve."""
    if grad not in ("envelope", "implicit"):
        raise ValueError("grad must be 'envelope' or 'implicit'")
    if grad == "implicit" and value != "primal":
        raise ValueError("grad='implicit' differentiates the transport cost: use value='primal' "
                         "(for value='dual' the envelope gradient is already exact)")
    n, m = x.shape[0], y.shape[0]
    if a is None:
        a = torch.full((n,), 1.0 / n, dtype=torch.float32, device=x.device)
    if b is None:
        b = torch.full((m,), 1.0 / m, dtype=torch.float32, device=x.device)
    return OTLoss.apply(x, y, a, b, float(eps), int(max_iter), float(tol), value, grad)


class ApplyPlan(torch.autograd.Function):
    """Z = [diag(1/P1)] P V, plan from (C, f, g) treated as a constant; gradient flows to V."""

    @staticmethod
    def forward(ctx, C, f, g, eps, V, normalise):
        Vd = V.detach().float().contiguous()
        Z, rs = ops.apply_plan(C, f, g, eps, Vd, normalise=normalise, return_rowsum=True)
        ctx.save_for_backward(C, f, g, rs)
        ctx.eps, ctx.normalise = eps, normalise
        return Z.to(V.dtype)

    @staticmethod
    def backward(ctx, dZ):
        C, f, g, rs = ctx.saved_tensors
        dtype = dZ.dtype
        dZ = dZ.float().contiguous()
        if ctx.normalise:
            dZ = dZ / torch.where(rs == 0, torch.full_like(rs, 1e-30), rs)[:, None]
        dV = ops.apply_plan(C, f, g, ctx.eps, dZ, transpose=True)
        return None, None, None, None, dV.to(dtype), None


def apply_plan(C, f, g, eps, V, normalise=False):
    return ApplyPlan.apply(C, f, g, float(eps), V, bool(normalise))
