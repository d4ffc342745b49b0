"""The reference's OT call surface, served by the B200 engine.

Each function keeps the name, argument meaning, return convention and warning
behaviour of the reference call it replaces (file:line relative to the reference
repository), so a training script swaps one import:

    ot.sinkhorn(a, b, M, reg, numItermax=..)        MRI_PET_OT_nojax.py:143
    sinkhorn_scaling(a, b, K, ...)                   perturbot/perturbot/match/utils.py:6-115
    linear_solve(Geometry(cost_matrix=..))           perturbot/perturbot/match/fot.py:129-134
    get_feature_coupling_pot(data, Ts, eps)          MRI_PET_OT_nojax.py:91-145
    fot_numpy / get_coupling_fot(data, Ts, eps)      perturbot/perturbot/match/fot.py:14-220
    mdict_to_matrix(M_dict, src, tgt)                baseline_models_fusion.py:233-239
    init_matrix_np(X1, X2, v1, v2)                   perturbot/perturbot/match/utils.py:125-184
    get_coupling_egw_ott_fixed(data, eps, ...)       MRI_PET_OT_OT_per_epoch_attn.py:129-186
    cotl_numpy / get_coupling_cotl_sinkhorn          perturbot/perturbot/match/cot_labels.py:14-341

Inputs may be NumPy arrays (as in the reference: copied to the GPU, result copied
back as NumPy in the input dtype) or torch tensors (CPU: same; CUDA: everything
stays on the device and the device->host->device round trip of
MRI_PET_OT_nojax.py:683-702 disappears).
"""
from __future__ import annotations

import math
import time
import warnings
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import B200OTError


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise B200OTError("b200ot needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class _Conv:
    """Remembers how the caller passed arrays so results go back the same way."""

    def __init__(self, ref, device=None):
        self.is_numpy = isinstance(ref, np.ndarray) or not isinstance(ref, torch.Tensor)
        if self.is_numpy:
            ref = np.asarray(ref)
            self.np_dtype = ref.dtype if ref.dtype.kind == "f" else np.dtype(np.float64)
            self.device = _device(device)
            self.home = None
        else:
            self.torch_dtype = ref.dtype if ref.dtype.is_floating_point else torch.float64
            self.home = ref.device
            self.device = ref.device if ref.is_cuda else _device(device)

    def to_dev(self, x, dtype=torch.float32) -> Optional[torch.Tensor]:
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=self.device, dtype=dtype, non_blocking=True)
        return torch.as_tensor(np.ascontiguousarray(x)).to(device=self.device, dtype=dtype, non_blocking=True)

    def back(self, t: torch.Tensor):
        if self.is_numpy:
            return t.detach().cpu().numpy().astype(self.np_dtype, copy=False)
        return t.to(device=self.home, dtype=self.torch_dtype)


_MASKED_COST = 1e30  # cost of a forbidden pair: exp(-1e30 / eps) is exactly 0 in fp32, and k * 1e30 stays finite


def _uniform(n, device):
    return torch.full((n,), 1.0 / n, dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------
# POT surface
# ---------------------------------------------------------------------------
def unif(n, type_as=None):
    """``ot.unif(n)`` (MRI_PET_OT_nojax.py:72-73): uniform histogram."""
    if isinstance(type_as, torch.Tensor):
        return torch.full((n,), 1.0 / n, dtype=type_as.dtype, device=type_as.device)
    return np.ones((n,)) / n


def dist(x1, x2=None, metric="sqeuclidean", *, device=None):
    """``ot.dist(x1, x2)`` (MRI_PET_OT_nojax.py:70-71): pairwise squared-Euclidean (default) or
    cosine cost, built on the GPU (tcgen05 split-bf16 GEMM for large problems)."""
    if metric not in ("sqeuclidean", "cosine"):
        raise B200OTError(f"metric {metric!r} is not used on the reference's OT path")
    x2 = x1 if x2 is None else x2
    cv = _Conv(x1, device)
    return cv.back(ops.cost_matrix(cv.to_dev(x1), cv.to_dev(x2), kind=metric))


def sinkhorn(a, b, M, reg, method="sinkhorn", numItermax=1000, stopThr=1e-9, verbose=False,
             log=False, warn=True, warmstart=None, *, check_every=10, err_norm="l2", stop_inclusive=False,
             path="auto", floor_patience=3, device=None, **kwargs):
    """Drop-in for ``ot.sinkhorn`` as called at MRI_PET_OT_nojax.py:143.

    Same stopping rule as POT 0.9.6 ``sinkhorn_knopp`` (L2 norm of the column-marginal
    violation checked when ``ii % 10 == 0``, stop when ``err < stopThr``), same start
    (``u = 1/n, v = 1/m``), same return value ``diag(u) K diag(v)`` -- evaluated in the
    log domain in fp32, so it stays finite where the reference's ``exp(-M/reg)``
    underflows.  Empty ``a`` / ``b`` mean uniform marginals, as in POT.

    ``floor_patience`` (default 3, 0 = off): fp32 cannot resolve POT's default ``stopThr = 1e-9`` on small
    problems (the marginal error bottoms out near 1e-7 |b|); once the error has set no new minimum for that
    many checks the solve stops instead of spinning to ``numItermax`` (``log["status"] == 1``).  Iteration
    counts equal the float64 reference's whenever ``stopThr`` is above that floor.
    """
    if method.lower() not in ("sinkhorn", "sinkhorn_log", "sinkhorn_stabilized"):
        raise B200OTError(f"method {method!r} is not part of the reference's OT path")
    cv = _Conv(M, device)
    Md = ops.aligned_copy(cv.to_dev(M))
    n, m = Md.shape
    ad = _uniform(n, cv.device) if a is None or len(a) == 0 else cv.to_dev(a)
    bd = _uniform(m, cv.device) if b is None or len(b) == 0 else cv.to_dev(b)
    if warmstart is not None:
        f0, g0 = (cv.to_dev(w) * float(reg) for w in warmstart)  # POT warmstart = (log u, log v)
    else:
        f0 = torch.full((n,), float(reg) * math.log(1.0 / n), dtype=torch.float32, device=cv.device)
        g0 = torch.full((m,), float(reg) * math.log(1.0 / m), dtype=torch.float32, device=cv.device)
    f, g, info = ops.sinkhorn_potentials(Md, ad, bd, float(reg), max_iter=int(numItermax),
                                         tol=float(stopThr), check_every=check_every, check_phase=1,
                                         err_norm=err_norm, stop_inclusive=stop_inclusive, path=path, f0=f0,
                                         g0=g0, floor_patience=floor_patience)
    if warn and not info["converged"] and stopThr > 0:
        warnings.warn("Sinkhorn did not converge. You might want to increase the number of "
                      "iterations `numItermax` or the regularization parameter `reg`.")
    P = ops.plan(Md, f, g, float(reg))
    out = cv.back(P)
    if not log:
        return out
    lg = {"err": cv.back(info["errs"]).tolist(), "niter": max(info["n_iter"] - 1, 0),
          "n_iter": info["n_iter"], "converged": info["converged"],
          "u": cv.back(torch.exp(f / float(reg))), "v": cv.back(torch.exp(g / float(reg))),
          "log_u": cv.back(f / float(reg)), "log_v": cv.back(g / float(reg)),
          "f": cv.back(f), "g": cv.back(g), "time": info["time"], "status": info["status"]}
    return out, lg


def sinkhorn_scaling(a, b, K, numItermax=1000, stopThr=1e-9, verbose=False, log=False,
                     always_raise=False, *, device=None, **kwargs):
    """Drop-in for the in-tree mirror ``sinkhorn_scaling`` (perturbot/perturbot/match/utils.py:6-115):
    takes the Gibbs kernel ``K``, squared-L2 error, ``while err > stopThr`` (inclusive stop).
    The engine works on ``-log K`` with reg = 1."""
    cv = _Conv(K, device)
    Kd = cv.to_dev(K, dtype=torch.float64)
    M = (-torch.log(Kd)).to(torch.float32)
    n, m = M.shape
    M = ops.aligned_copy(M)
    ad = _uniform(n, cv.device) if a is None or len(a) == 0 else cv.to_dev(a)
    bd = _uniform(m, cv.device) if b is None or len(b) == 0 else cv.to_dev(b)
    f0 = torch.full((n,), math.log(1.0 / n), dtype=torch.float32, device=cv.device)
    f, g, info = ops.sinkhorn_potentials(M, ad, bd, 1.0, max_iter=int(numItermax), tol=float(stopThr),
                                         check_every=10, check_phase=1, err_norm="l2sq",
                                         stop_inclusive=True, f0=f0)
    P = ops.plan(M, f, g, 1.0)
    out = cv.back(P)
    if not log:
        return out
    return out, {"err": cv.back(info["errs"]).tolist(), "u": cv.back(torch.exp(f)),
                 "v": cv.back(torch.exp(g)), "n_iter": info["n_iter"]}


# ---------------------------------------------------------------------------
# ott surface (the two objects the reference touches: Geometry and linear.solve)
# ---------------------------------------------------------------------------
class Geometry:
    """``ott.geometry.geometry.Geometry(cost_matrix=M, epsilon=eps, scale_cost="max_cost")``
    as constructed at perturbot/perturbot/match/fot.py:129-133."""

    def __init__(self, cost_matrix, epsilon=None, scale_cost=1.0, **kwargs):
        self.cost_matrix = cost_matrix
        self.epsilon = 0.05 if epsilon is None else float(epsilon)
        self.scale_cost = scale_cost


class SinkhornOutput:
    def __init__(self, matrix, f, g, n_iters, converged, errors, reg_ot_cost=None):
        self.matrix = matrix
        self.f, self.g = f, g
        self.n_iters = n_iters
        self.converged = converged
        self.errors = errors
        self.reg_ot_cost = reg_ot_cost


def linear_solve(geom: Geometry, a=None, b=None, max_iterations=2000, threshold=1e-3,
                 inner_iterations=10, *, path="auto", device=None, f0=None, g0=None, mask=None, _device_out=False,
                 **kwargs) -> SinkhornOutput:
    """Drop-in for ``ott.solvers.linear.solve(geom, max_iterations=N)`` (fot.py:129-134):
    cost divided by its max (``scale_cost="max_cost"``), eps absolute on the scaled cost,
    zero start potentials, g-then-f updates, L1 error of the b-marginal every
    ``inner_iterations``, stop when ``err < threshold``; ``.matrix`` is the plan."""
    cv = _Conv(geom.cost_matrix, device)
    Md = cv.to_dev(geom.cost_matrix)
    n, m = Md.shape
    Cs = ops.empty_matrix(n, m, cv.device)  # private copy: the scaling below is in place
    Cs.copy_(Md)
    if geom.scale_cost == "max_cost":
        ops.scale_by_inv_(Cs, ops.matrix_max(Cs))
    elif geom.scale_cost not in (None, 1.0, 1, "none"):
        raise B200OTError(f"scale_cost={geom.scale_cost!r} is not used by the reference")
    if mask is not None:  # plan restricted to the support of `mask` (label-aware solves): no mass elsewhere
        Cs.masked_fill_(~cv.to_dev(mask, dtype=torch.bool), _MASKED_COST)
    ad = _uniform(n, cv.device) if a is None else cv.to_dev(a)
    bd = _uniform(m, cv.device) if b is None else cv.to_dev(b)
    f, g, info = ops.sinkhorn_potentials(Cs, ad, bd, geom.epsilon, max_iter=int(max_iterations),
                                         tol=float(threshold), check_every=int(inner_iterations),
                                         check_phase=0, err_norm="l1", stop_inclusive=False, path=path,
                                         f0=None if f0 is None else cv.to_dev(f0),
                                         g0=None if g0 is None else cv.to_dev(g0))
    P = ops.plan(Cs, f, g, geom.epsilon)
    back = (lambda t: t) if _device_out else cv.back
    return SinkhornOutput(back(P), back(f), back(g), info["n_iter"], info["converged"], back(info["errs"]))


# ---------------------------------------------------------------------------
# helpers the reference callers use
# ---------------------------------------------------------------------------
def mdict_to_matrix(M_dict, source_labels, target_labels):
    """Block-diagonal scatter by label (baseline_models_fusion.py:233-239); host-side."""
    source_labels = np.asarray(source_labels)
    target_labels = np.asarray(target_labels)
    out = np.zeros((len(source_labels), len(target_labels)))
    for l, M in M_dict.items():
        out[np.ix_(np.where(source_labels == l)[0], np.where(target_labels == l)[0])] = np.asarray(M)
    return out


def init_matrix_np(X1, X2, v1, v2):
    """COOT square-loss factorisation (perturbot/perturbot/match/utils.py:125-184); host-side,
    trivial rank-1 sums.  The contraction it feeds runs on the GPU (ops.fot_cost)."""
    X1 = np.asarray(X1)
    X2 = np.asarray(X2)
    c1 = np.dot(X1 ** 2, np.asarray(v1, dtype=np.float64))
    c2 = np.dot(np.asarray(v2, dtype=np.float64), (X2 ** 2).T)
    return c1[:, None] + c2[None, :], X1, 2 * X2


def _block_diag_sorted(X_dict, Y_dict, Ts):
    keys = sorted(X_dict.keys())
    n_x = sum(len(X_dict[l]) for l in keys)
    n_y = sum(len(Y_dict[l]) for l in keys)
    out = np.zeros((n_x, n_y))
    ix = iy = 0
    for l in keys:
        nx, ny = len(X_dict[l]), len(Y_dict[l])
        if l in Ts:
            out[ix:ix + nx, iy:iy + ny] = np.asarray(Ts[l])
        ix += nx
        iy += ny
    return out


def _concat(d, keys):
    vals = [d[l] for l in keys]
    if isinstance(vals[0], torch.Tensor):
        return torch.cat(vals)
    return np.concatenate([np.asarray(v) for v in vals])


def get_feature_coupling_pot(data, Ts, eps=5e-3, *, numItermax=2000, stopThr=1e-9, err_norm="l2",
                             stop_inclusive=False, path="auto", device=None):
    """Drop-in for ``get_feature_coupling_pot`` (MRI_PET_OT_nojax.py:91-145): sorted-label concat,
    block-diagonal ``Ts``, feature cost ``M = t1 (+) t2 - 2 X^T Ts Y`` with ``w1 = Ts.sum(1)``,
    ``w2 = Ts.sum(0)``, uniform feature marginals, ``ot.sinkhorn(a, b, M, reg=eps,
    numItermax=2000)``.  Returns ``(Tv, {})``.  ``err_norm="l2"`` is POT 0.9.6's rule; the in-tree mirror
    of that loop (perturbot/perturbot/match/utils.py:88-89) uses ``err_norm="l2sq", stop_inclusive=True``."""
    X_dict, Y_dict = data
    keys = sorted(X_dict.keys())
    X = _concat(X_dict, keys)
    Y = _concat(Y_dict, keys)
    if isinstance(Ts, dict):
        Ts = _block_diag_sorted(X_dict, Y_dict, Ts)
    cv = _Conv(X, device)
    if cv.is_numpy:
        cv.np_dtype = np.dtype(np.float64)  # the reference's Tv is float64 (Ts promotes the cost)
    Xd, Yd, Td = cv.to_dev(X), cv.to_dev(Y), cv.to_dev(Ts)
    M = ops.fot_cost(Xd, Yd, Td, Td.sum(dim=1), Td.sum(dim=0))
    d1, d2 = M.shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Tv = sinkhorn(_uniform(d1, cv.device), _uniform(d2, cv.device), M, eps, numItermax=numItermax,
                      stopThr=stopThr, err_norm=err_norm, stop_inclusive=stop_inclusive, path=path)
    return cv.back(Tv), {}


def per_step_feature_plan(mri_feat, pet_feat, eps=1e-2, *, numItermax=2000, stopThr=1e-9, err_norm="l2",
                          stop_inclusive=False, device=None):
    """The per-step OT block of ``MultimodalMRI_PET_OT.forward`` (MRI_PET_OT_nojax.py:679-715) without the
    device -> host -> device round trip: identity sample coupling ``Ts = I / B`` (:687), ``get_feature_coupling_pot
    (({0: mri}, {0: pet}), {0: Ts}, eps=1e-2)`` (:694-698), then the guard ``NaN -> 1e-8`` and the row
    normalisation with zero sums replaced by 1e-8 (:704-715) fused into the kernel that writes the plan
    (``b200ot_plan_guard_rownorm``: the unnormalised plan is never stored).  Returns the (d_mri, d_pet) fp32
    coupling the forward applies as ``pet @ T.t()`` (:718); CUDA tensors in -> CUDA tensor out."""
    cv = _Conv(mri_feat, device)
    Xd, Yd = cv.to_dev(mri_feat).detach(), cv.to_dev(pet_feat).detach()
    B = Xd.shape[0]
    if Yd.shape[0] != B:
        raise B200OTError("per_step_feature_plan: one PET row per MRI row (identity sample coupling)")
    Ts = torch.eye(B, dtype=torch.float32, device=cv.device) / B
    w = torch.full((B,), 1.0 / B, dtype=torch.float32, device=cv.device)
    M = ops.fot_cost(Xd, Yd, Ts, w, w)
    d1, d2 = M.shape
    f0 = torch.full((d1,), float(eps) * math.log(1.0 / d1), dtype=torch.float32, device=cv.device)
    g0 = torch.full((d2,), float(eps) * math.log(1.0 / d2), dtype=torch.float32, device=cv.device)
    f, g, _ = ops.sinkhorn_potentials(M, _uniform(d1, cv.device), _uniform(d2, cv.device), float(eps),
                                      max_iter=int(numItermax), tol=float(stopThr), check_every=10, check_phase=1,
                                      err_norm=err_norm, stop_inclusive=stop_inclusive, f0=f0, g0=g0, floor_patience=3)
    T = ops.plan_guard_rownorm(M, f, g, float(eps))
    return T if not cv.is_numpy else cv.back(T)


def fot_numpy(X1, X2, Ts, v1=None, v2=None, niter=10, algo="emd", reg=0, algo2="emd", reg2=0, verbose=True,
              log=False, random_init=False, C_lin=None, *, device=None, path="auto", warm_start=None):
    """Drop-in for ``fot_numpy`` (perturbot/perturbot/match/fot.py:14-152), same positional order and defaults.
    With ``Ts`` fixed: ``Ts /= Ts.sum()``, ``w1 = Ts.sum(0)``, ``w2 = Ts.sum(1)`` (the reference's swapped axes,
    :109-110), cost ``constC - hC1 Ts hC2^T`` (:128), inner solve = ott ``linear.solve(Geometry(cost_matrix=M,
    epsilon=reg2, scale_cost="max_cost"), max_iterations=2000)`` (:129-134), ``cost = sum(M * Tv)`` (:137).

    As in the reference, ``v1``, ``v2``, ``algo``, ``algo2``, ``reg`` and ``C_lin`` do not enter the result (the
    feature solve is always the entropic ott solve with ``reg2``; ``v1``/``v2`` only seed ``random_init``).
    ``random_init=True`` needs scipy's sparse gamma initialisation of the reference and is rejected; ``reg2`` must be
    positive (the reference's default 0 makes ott's epsilon 0).  Because ``Ts`` never changes, the reference's
    second BCD round repeats the first solve bit for bit and exits on ``delta < 1e-16`` (:145): that repeat is not
    re-run, its cost is logged twice.  Returns ``(Tv, cost)`` or ``(Tv, cost, log)`` like the reference.

    ``warm_start=(f0, g0)`` (opt-in, not in the reference): start the solve from the potentials of a previous
    call (``log["potentials"]``), e.g. the previous epoch's; changes the iteration count, not the fixed point."""
    if random_init:
        raise B200OTError("fot_numpy: random_init needs scipy's sparse gamma initialisation and is not served")
    if not (reg2 > 0):
        raise B200OTError("fot_numpy: reg2 must be > 0 (it is ott's epsilon for the feature solve, fot.py:131)")
    cv = _Conv(X1, device)
    Xd, Yd = cv.to_dev(X1), cv.to_dev(X2)
    Td = cv.to_dev(Ts, dtype=torch.float64)
    Td = (Td / Td.sum()).to(torch.float32)
    M = ops.fot_cost(Xd, Yd, Td, Td.sum(dim=0), Td.sum(dim=1))
    cost_unscaled = M.clone()
    t0 = time.time()
    f0 = g0 = None
    if warm_start is not None:
        f0, g0 = (None if w is None else cv.to_dev(w) for w in warm_start)
    out = linear_solve(Geometry(cost_matrix=M, epsilon=reg2, scale_cost="max_cost"), max_iterations=2000,
                       path=path, f0=f0, g0=g0, _device_out=True)
    Tv_d = out.matrix
    cost = float((cost_unscaled.double() * Tv_d.double()).sum())
    rounds = [cost] if int(niter) <= 1 else [cost, cost]
    if verbose:
        for i, c in enumerate(rounds):
            print("Delta: {0}  Loss: {1}".format(0.0 if i else float(torch.linalg.norm(
                (Tv_d - 1.0 / Tv_d.numel()).double())), c))
        if len(rounds) > 1:
            print("converged at iter ", 1)
    if log:
        return cv.back(Tv_d), cost, {"cost": rounds, "time": time.time() - t0, "n_iters": out.n_iters,
                                     "potentials": (out.f, out.g)}
    return cv.back(Tv_d), cost


def get_coupling_fot(data, Ts, eps=5e-3, *, device=None, path="auto"):
    """Drop-in for ``get_coupling_fot`` (perturbot/perturbot/match/fot.py:155-220): first-seen
    label order, ``mdict_to_matrix`` block-diagonal ``Ts``, ``fot_numpy(X, Y, Ts, eps, eps, niter=2000)``.
    Returns ``(Tv, log)``."""
    X_dict, Y_dict = data
    keys = list(X_dict.keys())
    if isinstance(Ts, dict) and any(isinstance(v, torch.Tensor) and v.is_cuda for v in Ts.values()):
        # device-resident caller: the same block-diagonal scatter (rows / columns in first-seen label order are
        # contiguous blocks) without leaving the GPU
        dev = next(v.device for v in Ts.values() if isinstance(v, torch.Tensor) and v.is_cuda)
        nx = [len(X_dict[l]) for l in keys]
        ny = [len(Y_dict[l]) for l in keys]
        blk = torch.zeros((sum(nx), sum(ny)), dtype=torch.float64, device=dev)
        ix = iy = 0
        for l, a_, b_ in zip(keys, nx, ny):
            if l in Ts:
                blk[ix:ix + a_, iy:iy + b_] = torch.as_tensor(Ts[l], device=dev).to(torch.float64)
            ix += a_
            iy += b_
        Ts = blk
    elif isinstance(Ts, dict):
        Ts = mdict_to_matrix(
            Ts,
            np.concatenate([np.ones(len(X_dict[l])) * l for l in keys]),
            np.concatenate([np.ones(len(Y_dict[l])) * l for l in keys]))
    X = _concat(X_dict, keys)
    Y = _concat(Y_dict, keys)
    t0 = time.time()
    Tv, cost, lg = fot_numpy(X, Y, Ts, reg=eps, reg2=eps, niter=2000, log=True, verbose=False, device=device,
                             path=path)
    lg["time"] = time.time() - t0
    return Tv, lg


def group_features_by_label(y, p, max_samples_per_label=None):
    """Drop-in for ``group_features_by_label`` (MRI_PET_OT_OT_per_epoch_attn.py:918-937): bucket the rows of
    ``p`` by label (labels in ascending order, rows in their original order, at most ``max_samples_per_label``
    per bucket).  With CUDA tensors the buckets are built by device indexing: no device->host copy."""
    if isinstance(p, torch.Tensor):
        # one stable sort by label, one gather, ONE device -> host read (labels and bucket sizes together); the
        # buckets are views of the sorted copy.  (The first version synchronised once per label.)
        y = torch.as_tensor(y, device=p.device).reshape(-1)
        order = torch.argsort(y, stable=True)
        labels, counts = torch.unique_consecutive(y.index_select(0, order), return_counts=True)
        sorted_p = p.index_select(0, order)
        meta = torch.stack([labels.to(torch.int64), counts.to(torch.int64)]).tolist()
        out, start = {}, 0
        for label, cnt in zip(*meta):
            take = cnt if not (max_samples_per_label is not None and max_samples_per_label > 0) else min(
                cnt, int(max_samples_per_label))
            out[int(label)] = sorted_p[start:start + take]
            start += cnt
        return out
    y = np.asarray(y)
    p = np.asarray(p)
    out = {}
    for label in np.unique(y):
        arr = p[y == label]
        if max_samples_per_label is not None and max_samples_per_label > 0:
            arr = arr[:max_samples_per_label]
        out[int(label)] = arr
    return out


def cotl_numpy(X_dict, Y_dict, w1=None, w2=None, v1=None, v2=None, niter=100, algo="emd", reg=0.2, algo2="emd",
               reg2=0.2, verbose=True, log=False, random_init=False, C_lin=None, *, device=None, warm_start=False):
    """Drop-in for ``cotl_numpy`` (perturbot/perturbot/match/cot_labels.py:14-225) in its entropic form
    (``algo="sinkhorn", algo2="sinkhorn"``, what ``get_coupling_cotl_sinkhorn`` runs): label-constrained COOT by block
    coordinate descent.  Per label a sample coupling on ``constC_s - hC1_s Tv hC2_s^T`` (:172), then one feature
    coupling on the sum over labels of ``constC_v - hC1_v Ts_k hC2_v^T`` (:190-193); every cost is built on the GPU
    (``b200ot_fot_cost``) and every solve is the engine's ott-flavoured ``linear_solve``; couplings stay on the
    device between rounds.  Reference quirks kept: the feature solve uses ``reg`` (``reg2`` is never read, :201);
    ``Tsold = Ts`` aliases the dict so ``delta`` only sees ``Tv`` (:162,209-211); exit on ``delta < 1e-16`` or
    ``|cost_old - cost| < 1e-7`` (:219).  The ``"emd"`` variants are POT's network simplex and are not served.

    ``warm_start=True`` (opt-in, not in the reference): every solve of BCD round r + 1 starts from the
    potentials its counterpart reached in round r instead of from zero.  The costs change little between rounds, so
    the inner solves converge in a fraction of the iterations (``log["inner_iterations"]`` records the total); the
    fixed point of each solve is the same, the iterates (and so the last digits of the couplings) are not."""
    if algo != "sinkhorn" or algo2 != "sinkhorn":
        raise B200OTError("cotl_numpy: only algo='sinkhorn', algo2='sinkhorn' run on the B200 engine "
                          "('emd' is POT's network-simplex solver)")
    if random_init:
        raise B200OTError("cotl_numpy: random_init needs scipy's sparse gamma initialisation and is not served")
    labels = list(X_dict.keys())
    if sorted(labels) != sorted(Y_dict.keys()):
        raise AssertionError("Labels don't match in y1 & y2.")
    cv = _Conv(X_dict[labels[0]], device)
    dev = cv.device
    Xd = {k: cv.to_dev(X_dict[k]).contiguous() for k in labels}
    Yd = {k: cv.to_dev(Y_dict[k]).contiguous() for k in labels}
    Xt = {k: Xd[k].t().contiguous() for k in labels}
    Yt = {k: Yd[k].t().contiguous() for k in labels}
    d1, d2 = Xd[labels[0]].shape[1], Yd[labels[0]].shape[1]

    def feature_weights(v, parts, d):
        if v is not None:
            return cv.to_dev(v)
        allx = torch.cat([parts[k] for k in labels], dim=0).double()
        if bool((allx >= 0).all()):
            return (allx.sum(0) / allx.sum()).float()
        return torch.full((d,), 1.0 / d, dtype=torch.float32, device=dev)

    v1d, v2d = feature_weights(v1, Xd, d1), feature_weights(v2, Yd, d2)
    w1d = {k: (cv.to_dev(w1[k]) if w1 is not None else _uniform(Xd[k].shape[0], dev)) for k in labels}
    w2d = {k: (cv.to_dev(w2[k]) if w2 is not None else _uniform(Yd[k].shape[0], dev)) for k in labels}
    Clin = None if C_lin is None else cv.to_dev(C_lin)
    Ts = {k: torch.full((Xd[k].shape[0], Yd[k].shape[0]), 1.0 / (Xd[k].shape[0] * Yd[k].shape[0]),
                        dtype=torch.float32, device=dev) for k in labels}
    Tv = torch.full((d1, d2), 1.0 / (d1 * d2), dtype=torch.float32, device=dev)
    cost = float("inf")
    log_out = {"cost": [], "inner_iterations": 0}
    pot = {}  # potentials of the previous round's solves (warm_start)
    for i in range(int(niter)):
        Tv_old, cost_old = Tv, cost
        for k in labels:  # sample OT per label: rows of the transposed data play the role of features
            M_k = ops.fot_cost(Xt[k], Yt[k], Tv, v1d, v2d)
            if Clin is not None:
                M_k = M_k + Clin
            f0, g0 = pot.get(k, (None, None)) if warm_start else (None, None)
            out_k = linear_solve(Geometry(cost_matrix=M_k, epsilon=reg, scale_cost="max_cost"),
                                 max_iterations=2000, f0=f0, g0=g0, _device_out=True)
            Ts[k], pot[k] = out_k.matrix, (out_k.f, out_k.g)
            log_out["inner_iterations"] += out_k.n_iters
        M = None
        for k in labels:  # global feature OT on the summed cost
            Mk = ops.fot_cost(Xd[k], Yd[k], Ts[k], w1d[k], w2d[k])
            M = Mk if M is None else M + Mk
        f0, g0 = pot.get("_features", (None, None)) if warm_start else (None, None)
        out_v = linear_solve(Geometry(cost_matrix=M, epsilon=reg, scale_cost="max_cost"), max_iterations=2000,
                             f0=f0, g0=g0, _device_out=True)
        Tv, pot["_features"] = out_v.matrix, (out_v.f, out_v.g)
        log_out["inner_iterations"] += out_v.n_iters
        tot = float(Tv.double().sum())
        if not abs(tot - 1.0) < 1e-8:
            Tv = (Tv.double() / tot).float()
        delta = float(torch.linalg.norm((Tv - Tv_old).double()))
        cost = float((M.double() * Tv.double()).sum())
        if log:
            log_out["cost"].append(cost)
        if verbose:
            print(f"It {i} Delta: {delta}  Loss: {cost}")
        if delta < 1e-16 or abs(cost_old - cost) < 1e-7:
            if verbose:
                print("converged at iter ", i)
            break
    Ts_out = {k: cv.back(Ts[k]) for k in labels}
    if log:
        return Ts_out, cv.back(Tv), cost, log_out
    return Ts_out, cv.back(Tv), cost


def get_coupling_cotl_sinkhorn(data, eps: float = 5e-3, eps2: Optional[float] = None, *, device=None):
    """Drop-in for ``get_coupling_cotl_sinkhorn`` (perturbot/perturbot/match/cot_labels.py:283-341):
    ``(Ts per label, log)``, or ``(-1, -1)`` on a floating-point failure like the reference (:337-338)."""
    X_dict, Y_dict = data
    start = time.time()
    if eps2 is None:
        eps2 = eps
    try:
        Ts, Tv, cost, log = cotl_numpy(X_dict, Y_dict, algo="sinkhorn", reg=eps, algo2="sinkhorn", reg2=eps2,
                                       log=True, niter=2000, verbose=False, device=device)
    except FloatingPointError:
        return -1, -1
    log["time"] = time.time() - start
    return Ts, log


def get_coupling_egw_ott_fixed(data, eps: float = 5e-3, gw_max_iterations: int = 2000,
                               sinkhorn_max_iterations: int = 2000, *, device=None):
    """Drop-in for ``get_coupling_egw_ott_fixed`` (MRI_PET_OT_OT_per_epoch_attn.py:129-186; also
    ``get_coupling_egw_ott``, MRI_PET_OT.py:68-122): entropic Gromov-Wasserstein sample coupling per label between
    the MRI and PET embeddings of that label, on max-scaled squared-Euclidean point-cloud geometries.  All labels
    are solved by ONE kernel launch, one CTA per label.  ``data = (X_dict, Y_dict)`` with NumPy arrays (as in the
    reference: results come back as NumPy) or torch tensors (CUDA: couplings stay on the device).  Returns
    ``(Ts, log)`` with the reference's log keys; NaN features are mapped to 0 with a message (:148-151).
    Labels with more than 64 samples on either side (``max_samples_per_label=None``; ``--max-jax-samples 128``) do
    not fit the shared-memory kernel and are solved one at a time on the dense path (``_egw_dense``, fp32)."""
    X_dict, Y_dict = data
    labels = list(X_dict.keys())
    conv = _Conv(X_dict[labels[0]], device)
    t0 = time.time()
    Xs, Ys = [], []
    for l in labels:
        x = conv.to_dev(X_dict[l])
        y = conv.to_dev(Y_dict[l])
        if bool(torch.isnan(x).any()) or bool(torch.isnan(y).any()):
            print(f"Warning: NaNs detected in features for label {l}")
            x, y = torch.nan_to_num(x), torch.nan_to_num(y)
        Xs.append(x)
        Ys.append(y)
    cost_time = time.time() - t0
    t0 = time.time()
    # labels of at most 64 x 64 samples: ONE launch, one CTA per label, everything in shared memory (float64);
    # larger labels (max_samples_per_label=None, or args.max_jax_samples = 128) take the dense path, one at a time
    small = [i for i in range(len(labels)) if Xs[i].shape[0] <= _EGW_SMEM_MAX and Ys[i].shape[0] <= _EGW_SMEM_MAX]
    out_T, log = {}, {}
    if small:
        Ts, info = ops.egw_batched([Xs[i] for i in small], [Ys[i] for i in small], eps, gw_max_iterations,
                                   sinkhorn_max_iterations)
        host = {k: v.cpu() for k, v in info.items()}
        for j, i in enumerate(small):
            out_T[labels[i]] = conv.back(Ts[j])
            log[labels[i]] = {"n_iters_outer": int(host["n_iters_outer"][j]),
                              "converged_inner": bool(host["converged_inner"][j]),
                              "converged_outer": bool(host["converged_outer"][j]), "GW cost": float(host["GW cost"][j]),
                              "inner_iterations": int(host["inner_iterations"][j])}
    for i in range(len(labels)):
        if i in small:
            continue
        T, lg = _egw_dense(Xs[i], Ys[i], eps, gw_max_iterations, sinkhorn_max_iterations)
        out_T[labels[i]] = conv.back(T)
        log[labels[i]] = lg
    dt = time.time() - t0
    out_T = {l: out_T[l] for l in labels}  # the reference's key order
    for l in labels:
        log[l]["time"], log[l]["cost_time"] = dt, cost_time
    return out_T, {l: log[l] for l in labels}


_EGW_SMEM_MAX = 64  # kEgwMax of csrc/egw.cu: samples per side the one-CTA-per-label kernel holds in shared memory


def _egw_dense(X, Y, eps, gw_max_iterations=2000, sinkhorn_max_iterations=2000, mask=None, gw_threshold=1e-3,
               gw_min_iterations=5, sk_threshold=1e-3, sk_check_every=10):
    """Entropic Gromov-Wasserstein of ANY size on device tensors: ott ``GromovWasserstein`` semantics as in
    ``get_coupling_egw_ott_fixed`` (squared-Euclidean geometries divided by their maximum, square loss, uniform
    marginals, warm-started inner log-domain Sinkhorn, outer stop on ``isclose(cost[-2], cost[-1], rtol=1e-3)`` after
    5 iterations), with the per-iteration work on the engine's kernels: the linearised cost
    ``(C1^2) T1 (+) (C2^2) T^T 1 - 2 C1 T C2`` is ``b200ot_fot_cost``, the inner solve ``b200ot_sinkhorn_solve``
    (resident kernel for these sizes), the coupling ``b200ot_plan``.  ``mask`` restricts the coupling to its
    support (label-aware form).  fp32, unlike the float64 shared-memory kernel for labels of <= 64 samples."""
    dev = X.device
    n, m = X.shape[0], Y.shape[0]
    C1 = ops.cost_matrix(X, X, impl="simt" if n < 512 else "auto")
    C2 = ops.cost_matrix(Y, Y, impl="simt" if m < 512 else "auto")
    C1.clamp_(min=0.0)  # |x_i - x_i|^2 is 0, not -1e-7
    C2.clamp_(min=0.0)
    ops.scale_by_inv_(C1, ops.matrix_max(C1))
    ops.scale_by_inv_(C2, ops.matrix_max(C2))
    a, b = _uniform(n, dev), _uniform(m, dev)
    T = torch.outer(a, b)
    if mask is not None:
        mask = mask.to(device=dev, dtype=torch.bool)
        T = T * mask
    f = torch.zeros(n, dtype=torch.float32, device=dev)
    g = torch.zeros(m, dtype=torch.float32, device=dev)
    costs, inner_total, inner_conv, outer_conv = [], 0, False, False
    while len(costs) < int(gw_max_iterations):
        M = ops.fot_cost(C1, C2, T, T.sum(dim=1), T.sum(dim=0))
        if mask is not None:
            M.masked_fill_(~mask, _MASKED_COST)
        f, g, info = ops.sinkhorn_potentials(M, a, b, float(eps), max_iter=int(sinkhorn_max_iterations),
                                             tol=float(sk_threshold), check_every=int(sk_check_every), check_phase=0,
                                             err_norm="l1", f0=f, g0=g)
        inner_total += info["n_iter"]
        inner_conv = bool(info["converged"])
        T = ops.plan(M, f, g, float(eps))
        ff, gg = f.double(), g.double()
        costs.append(float(ff[torch.isfinite(ff)].sum() / n + gg[torch.isfinite(gg)].sum() / m))
        if not math.isfinite(costs[-1]):
            break
        k = len(costs)
        if k >= 2:
            outer_conv = abs(costs[-2] - costs[-1]) <= 1e-8 + gw_threshold * abs(costs[-1])
            if outer_conv and k >= gw_min_iterations:
                break
    return T, {"n_iters_outer": len(costs), "converged_inner": inner_conv, "converged_outer": bool(outer_conv),
               "GW cost": costs[-1], "inner_iterations": inner_total}


def _concat_with_labels(data, cv):
    """Concatenate both sides in first-seen label order with their label vectors (ott_egwl.py:66-75)."""
    X_dict, Y_dict = data
    keys = list(X_dict.keys())
    X = torch.cat([cv.to_dev(X_dict[l]) for l in keys])
    Y = torch.cat([cv.to_dev(Y_dict[l]) for l in keys])
    sl = np.concatenate([np.repeat(l, len(X_dict[l])) for l in keys])
    tl = np.concatenate([np.repeat(l, len(Y_dict[l])) for l in keys])
    return X, Y, sl, tl


def _label_mask(sl, tl, device):
    """``create_block_diag_mat`` (ott_egwl.py:16-22) as a boolean device tensor: True where the labels agree."""
    s_t = torch.as_tensor(np.unique(sl, return_inverse=True)[1], device=device)
    uniq = {v: i for i, v in enumerate(np.unique(sl))}
    t_t = torch.as_tensor(np.array([uniq.get(v, -1) for v in tl]), device=device)
    return s_t[:, None] == t_t[None, :]


def _blocks_by_label(T, sl, tl, cv):
    dev = T.device
    out = {}
    for l in np.unique(sl):
        ri = torch.as_tensor(np.where(sl == l)[0], device=dev)
        ci = torch.as_tensor(np.where(tl == l)[0], device=dev)
        out[l] = cv.back(T.index_select(0, ri).index_select(1, ci))
    return out


def get_coupling_egw_all_ott(data, eps: float = 5e-3, *, device=None):
    """Drop-in for ``get_coupling_egw_all_ott`` (perturbot/perturbot/match/ott_egwl.py:209-297): ONE entropic
    Gromov-Wasserstein problem over all samples of all labels (labels disregarded), ``GromovWasserstein(epsilon=eps,
    max_iterations=1000)`` with the default inner ``Sinkhorn``.  Returns ``(T, log)`` with the full (n, m) coupling."""
    cv = _Conv(next(iter(data[0].values())), device)
    t0 = time.time()
    X, Y, _, _ = _concat_with_labels(data, cv)
    cost_time = time.time() - t0
    t0 = time.time()
    T, log = _egw_dense(X, Y, eps, gw_max_iterations=1000, sinkhorn_max_iterations=2000)
    log["time"], log["cost_time"] = time.time() - t0, cost_time
    return cv.back(T), log


def get_coupling_egw_labels_ott(data, eps: float = 5e-3, *, device=None):
    """Drop-in for ``get_coupling_egw_labels_ott`` (perturbot/perturbot/match/ott_egwl.py:25-127): entropic
    Gromov-Wasserstein over all samples with the coupling restricted to pairs of equal label
    (``T_ij > 0 => l_{x_i} = l_{y_j}``, :37; ``block_diag_mat``, :16-22,84), ``max_iterations=2000`` outer and inner.
    The reference calls a modified OTT (``labels_a / labels_b / n_labels / block_diag_mat`` on ``QuadraticProblem``)
    that is not in its tree; the constraint is served as a support mask on the start coupling and on every
    linearised cost, global uniform marginals kept.  Returns ``(T_dict, log)``: the diagonal blocks keyed by label
    in ``np.unique`` order, like the reference (:124-127)."""
    cv = _Conv(next(iter(data[0].values())), device)
    t0 = time.time()
    X, Y, sl, tl = _concat_with_labels(data, cv)
    mask = _label_mask(sl, tl, X.device)
    cost_time = time.time() - t0
    t0 = time.time()
    T, log = _egw_dense(X, Y, eps, gw_max_iterations=2000, sinkhorn_max_iterations=2000, mask=mask)
    log["time"], log["cost_time"] = time.time() - t0, cost_time
    return _blocks_by_label(T, sl, tl, cv), log


def get_coupling_leot_ott(data, eps: float = 5e-3, *, device=None):
    """Drop-in for ``get_coupling_leot_ott`` (perturbot/perturbot/match/ott_egwl.py:375-454): label-constrained
    entropic OT.  Squared-Euclidean cost over all samples divided by its maximum
    (``PointCloud(scale_cost="max_cost").cost_matrix``, :426-427; built on tcgen05 for large inputs), then ott
    ``Sinkhorn()`` on ``LinearProblem(geom, labels_a, labels_b)`` of the modified OTT: served as the engine's
    ott-flavoured solve with the plan restricted to pairs of equal label.  Returns ``(T_dict, log)`` with the
    reference's log keys (``OT cost`` = dual value)."""
    cv = _Conv(next(iter(data[0].values())), device)
    t0 = time.time()
    X, Y, sl, tl = _concat_with_labels(data, cv)
    C = ops.cost_matrix(X, Y)
    mask = _label_mask(sl, tl, X.device)
    cost_time = time.time() - t0
    t0 = time.time()
    out = linear_solve(Geometry(cost_matrix=C, epsilon=eps, scale_cost="max_cost"), mask=mask, _device_out=True)
    n, m = C.shape
    f, g = out.f.double(), out.g.double()
    log = {"n_iters_outer": out.n_iters, "converged": out.converged,
           "OT cost": float(f[torch.isfinite(f)].sum() / n + g[torch.isfinite(g)].sum() / m)}
    log["time"], log["cost_time"] = time.time() - t0, cost_time
    return _blocks_by_label(out.matrix, sl, tl, cv), log


def get_coupling_egw_ott(data, eps: float = 5e-3, *, device=None):
    """Drop-in for ``get_coupling_egw_ott`` (perturbot/perturbot/match/ott_egwl.py:129-206; MRI_PET_OT.py:68-122):
    the same per-label entropic Gromov-Wasserstein solve with ott's defaults spelled out there --
    ``GromovWasserstein(epsilon=eps, max_iterations=1000)`` and the default inner ``Sinkhorn()`` (2000 iterations)."""
    return get_coupling_egw_ott_fixed(data, eps, gw_max_iterations=1000, sinkhorn_max_iterations=2000, device=device)


def get_coupling_eot_ott(data, eps: float = 5e-3, *, device=None):
    """Drop-in for ``get_coupling_eot_ott`` (perturbot/perturbot/match/ott_egwl.py:299-372): ONE sample-level
    entropic OT problem over all labels concatenated (label information is disregarded): squared-Euclidean
    point-cloud cost divided by its maximum (``PointCloud(scale_cost="max_cost").cost_matrix``, :350), then
    ``linear.solve(Geometry(cost_matrix, epsilon=eps))`` (:356).  The cost is built on the tcgen05 kernel and the
    solve is the engine's ott-flavoured log-domain Sinkhorn.  Returns ``(T, log)`` with the reference's log keys
    (``OT cost`` is the dual value ``<a, f> + <b, g>``)."""
    X_dict, Y_dict = data
    keys = list(X_dict.keys())
    cv = _Conv(X_dict[keys[0]], device)
    t0 = time.time()
    X = torch.cat([cv.to_dev(X_dict[l]) for l in keys])
    Y = torch.cat([cv.to_dev(Y_dict[l]) for l in keys])
    C = ops.cost_matrix(X, Y)
    cost_time = time.time() - t0
    t0 = time.time()
    out = linear_solve(Geometry(cost_matrix=C, epsilon=eps, scale_cost="max_cost"))
    n, m = C.shape
    f, g = torch.as_tensor(out.f).double(), torch.as_tensor(out.g).double()
    log = {"n_iters_outer": out.n_iters, "converged": out.converged,
           "OT cost": float(f[torch.isfinite(f)].sum() / n + g[torch.isfinite(g)].sum() / m)}
    T = cv.back(out.matrix)
    log["time"] = time.time() - t0
    log["cost_time"] = cost_time
    return T, log


def compute_pet_to_mri_coupling(mri_features, pet_features, labels, max_samples_per_label=None,
                                gw_max_iterations: int = 2000, sinkhorn_max_iterations: int = 2000, *, device=None):
    """The OT part of ``compute_pet_to_mri_coupling`` (MRI_PET_OT_OT_per_epoch_attn.py:940-960), i.e. everything
    after ``feature_extract``: bucket both modalities by label (``max_samples_per_label`` = ``args.max_jax_samples``),
    entropic Gromov-Wasserstein sample coupling per label (PET -> MRI), feature coupling ``get_coupling_fot`` on
    the block-diagonal of those.  With CUDA tensors nothing leaves the device (the reference converts to NumPy,
    solves on the CPU and copies the d x d coupling back, :943-953,1233-1236).  Returns the (d_pet, d_mri)
    feature coupling ``T_feature_pet2mri``."""
    grouped_mri = group_features_by_label(labels, mri_features, max_samples_per_label=max_samples_per_label)
    grouped_pet = group_features_by_label(labels, pet_features, max_samples_per_label=max_samples_per_label)
    T_dict, _ = get_coupling_egw_ott_fixed((grouped_pet, grouped_mri), gw_max_iterations=gw_max_iterations,
                                           sinkhorn_max_iterations=sinkhorn_max_iterations, device=device)
    T_feature, _ = get_coupling_fot((grouped_pet, grouped_mri), T_dict, device=device)
    return T_feature


def foscttm(Y_pred, Y_true, idx=None, *, device=None):
    """Drop-in for ``foscttm`` (perturbot/perturbot/eval/utils.py:18-45): fraction of samples closer than the
    true match, per sample, as a list.  The O(n^2) distance matrix and the per-row ranks are computed on the GPU
    (the reference sorts one row at a time in Python)."""
    if idx is not None:
        Y_pred, Y_true = Y_pred[:, idx], Y_true[:, idx]
    cv = _Conv(Y_pred, device)
    out = ops.foscttm(cv.to_dev(Y_pred), cv.to_dev(Y_true))
    return out.double().cpu().tolist()


def get_FOSCTTM(T, Xs_true, Xt_true, use_barycenter=True, use_agg="mean", *, device=None):
    """Drop-in for the dense-plan branch of ``get_FOSCTTM`` (perturbot/perturbot/eval/match.py:178-206):
    barycentric projection ``(T / rowsum) @ Xt`` (``rowsum == 0 -> 1e-30``) then FOSCTTM, aggregated."""
    cv = _Conv(Xt_true, device)
    Xt = cv.to_dev(Xt_true)
    if use_barycenter:
        Td = cv.to_dev(T)
        marg = Td.sum(dim=-1, keepdim=True)
        marg = torch.where(marg == 0, torch.full_like(marg, 1e-30), marg)
        pred = (Td / marg) @ Xt
    else:
        pred = cv.to_dev(Xs_true)
    vals = ops.foscttm(pred.contiguous(), Xt).double().cpu().numpy()
    agg = np.nanmedian if use_agg == "median" else np.nanmean
    return vals.tolist(), float(agg(vals))


# ---------------------------------------------------------------------------
# north-star surface: embeddings in, plan / loss / fused embedding out
# ---------------------------------------------------------------------------
def sinkhorn_from_embeddings(x, y, a=None, b=None, reg=0.05, numItermax=1000, stopThr=1e-9,
                             cost="sqeuclidean", check_every=10, err_norm="l2", path="auto",
                             return_plan=False, V=None, device=None, fused_out=None):
    """Embeddings (host or device) -> cost on the GPU -> Sinkhorn -> potentials, OT cost and,
    if ``V`` is given, the barycentric projection ``diag(1/P1) P V`` (the fused embedding; single-pass
    tcgen05 kernel); the plan itself is only materialised when ``return_plan`` is set.  ``fused_out``: optional
    preallocated (e.g. pinned host) tensor that receives the fused embedding.  Returns a dict."""
    cv = _Conv(x, device)
    xd, yd = cv.to_dev(x), cv.to_dev(y)
    n, m = xd.shape[0], yd.shape[0]
    Cm = ops.cost_matrix(xd, yd, kind=cost)
    ad = _uniform(n, cv.device) if a is None else cv.to_dev(a)
    bd = _uniform(m, cv.device) if b is None else cv.to_dev(b)
    f0 = torch.full((n,), float(reg) * math.log(1.0 / n), dtype=torch.float32, device=cv.device)
    f, g, info = ops.sinkhorn_potentials(Cm, ad, bd, float(reg), max_iter=int(numItermax), tol=float(stopThr),
                                         check_every=check_every, check_phase=1, err_norm=err_norm,
                                         path=path, f0=f0)
    out = {"f": cv.back(f), "g": cv.back(g), "n_iter": info["n_iter"], "converged": info["converged"],
           "err": info["err"], "ot_cost": float(ops.ot_cost(Cm, f, g, float(reg)).item())}
    if V is not None:
        fused = ops.apply_plan(Cm, f, g, float(reg), cv.to_dev(V), normalise=True)
        if fused_out is not None:
            fused_out.copy_(fused, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out["fused"] = fused_out
        else:
            out["fused"] = cv.back(fused)
    if return_plan:
        out["plan"] = cv.back(ops.plan(Cm, f, g, float(reg)))
    return out
