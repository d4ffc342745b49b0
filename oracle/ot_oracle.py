"""CPU oracle for the entropic-OT (Sinkhorn) MRI<->PET alignment path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU arm -- never as a fallback for the CUDA path.

It restates, in float64 NumPy, the algorithm the reference runs on this path.
All ``file:line`` citations are relative to ``/root/reference``.

Parity status
-------------
* ``sinkhorn_knopp`` / ``init_matrix`` / ``fot_cost_pot`` / ``mdict_to_matrix``
  are PINNED: ``tests/golden/make_golden.py`` executes the reference's own code
  (``perturbot/perturbot/match/utils.py`` imported by file path, and the
  function bodies of ``MRI_PET_OT_nojax.py:91-145`` and
  ``perturbot/perturbot/match/fot.py:14-152`` compiled from the reference file
  with ``ot`` / ``ott`` replaced by stubs) and the resulting vectors are
  committed under ``tests/golden/``; ``tests/test_oracle.py`` checks the oracle
  against them.
* ``sinkhorn_log_ott`` restates ott-jax 0.6.0 ``linear.solve`` semantics
  (``OT_environment.yml:227``), and ``err_norm="l2"`` restates POT 0.9.6.post1
  (``OT_environment.yml:232``).  Neither library is vendored nor installable
  here, the reference ships no tests/golden vectors for them, so for these two
  flavours **parity is unpinned**: they are restatements of the published
  algorithms anchored on the reference's call sites
  (``MRI_PET_OT_nojax.py:143``, ``perturbot/perturbot/match/fot.py:129-134``).
* ``cotl_sinkhorn`` (BCD shell of ``cotl_numpy``, ``perturbot/perturbot/match/cot_labels.py:14-225``) is PINNED the
  same way (reference function compiled from its AST, inner ``linear.solve`` bound to ``sinkhorn_log_ott``).
* ``egw_ott`` restates ott-jax 0.6.0 ``GromovWasserstein`` (square loss, constant epsilon, warm-started inner
  Sinkhorn) as called at ``MRI_PET_OT_OT_per_epoch_attn.py:155-175``; same situation, **parity unpinned**.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "synthetic_embeddings", "sqeuclid_cost", "cosine_cost", "init_matrix",
    "fot_cost_pot", "fot_cost_ott", "mdict_to_matrix", "sinkhorn_knopp",
    "sinkhorn_log", "sinkhorn_log_ott", "plan_from_potentials", "rows_given_g", "fot_bcd_ott",
    "get_feature_coupling_pot", "get_coupling_fot", "plan_guard_rownorm",
    "apply_plan_T", "barycentric", "cosine_loss", "ot_cost", "envelope_grads", "foscttm",
    "group_features_by_label", "egw_ott", "get_coupling_egw_ott_fixed", "cotl_sinkhorn", "block_diag_mask",
    "get_coupling_egw_all_ott", "get_coupling_egw_labels_ott", "get_coupling_leot_ott",
]


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def synthetic_embeddings(n, m, d, config_index=0, seed_base=20251118):
    """X = randn(n,d); Y = randn(m,d) + 0.5*randn(1,d); rows L2-normalised, fp32.

    Uses torch's CPU generator so bench.py, the tests and the golden script all
    see the same bits ("Synthetic inputs").
    """
    import torch

    gen = torch.Generator(device="cpu").manual_seed(seed_base + config_index)
    X = torch.randn(n, d, generator=gen, dtype=torch.float32)
    Y = torch.randn(m, d, generator=gen, dtype=torch.float32)
    Y = Y + 0.5 * torch.randn(1, d, generator=gen, dtype=torch.float32)
    X = X / X.norm(dim=1, keepdim=True)
    Y = Y / Y.norm(dim=1, keepdim=True)
    return X.numpy(), Y.numpy()


# --------------------------------------------------------------------------
# cost construction
# --------------------------------------------------------------------------
def sqeuclid_cost(X, Y):
    """C_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  (north-star sample x sample cost).

    Same algebra as the reference's feature cost with Ts = I
    (``MRI_PET_OT_nojax.py:121-136``).
    """
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    return (X * X).sum(1)[:, None] + (Y * Y).sum(1)[None, :] - 2.0 * (X @ Y.T)


def cosine_cost(X, Y):
    """C_ij = 1 - cos(x_i, y_j): the same contraction on pre-normalised rows."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    Xn = X / np.maximum(np.linalg.norm(X, axis=1, keepdims=True), 1e-300)
    Yn = Y / np.maximum(np.linalg.norm(Y, axis=1, keepdims=True), 1e-300)
    return 1.0 - Xn @ Yn.T


def init_matrix(X1, X2, v1, v2):
    """COOT square-loss factorisation (``perturbot/perturbot/match/utils.py:125-184``).

    constC = (X1^2) v1 (+) (X2^2) v2 ; hC1 = X1 ; hC2 = 2 X2, so that the cost is
    ``constC - hC1 . T . hC2^T``.
    """
    # No up-cast of X1/X2: the reference squares in the input dtype (float32
    # embeddings) and only promotes to float64 in the product with the weights.
    X1 = np.asarray(X1)
    X2 = np.asarray(X2)
    c1 = np.dot(X1 ** 2, np.asarray(v1, dtype=np.float64))
    c2 = np.dot(np.asarray(v2, dtype=np.float64), (X2 ** 2).T)
    return c1[:, None] + c2[None, :], X1, 2 * X2


def mdict_to_matrix(M_dict, source_labels, target_labels):
    """Block-diagonal scatter by label (``baseline_models_fusion.py:233-239``)."""
    source_labels = np.asarray(source_labels)
    target_labels = np.asarray(target_labels)
    out = np.zeros((len(source_labels), len(target_labels)))
    for l, M in M_dict.items():
        out[np.ix_(np.where(source_labels == l)[0], np.where(target_labels == l)[0])] = M
    return out


def _ts_from_dict_sorted(X_dict, Y_dict, Ts):
    """Block-diagonal Ts in sorted-label order (``MRI_PET_OT_nojax.py:105-119``)."""
    keys = sorted(X_dict.keys())
    n_x = sum(len(X_dict[l]) for l in keys)
    n_y = sum(len(Y_dict[l]) for l in keys)
    out = np.zeros((n_x, n_y))
    ix = iy = 0
    for l in keys:
        nx, ny = len(X_dict[l]), len(Y_dict[l])
        if l in Ts:
            out[ix:ix + nx, iy:iy + ny] = Ts[l]
        ix += nx
        iy += ny
    return out


def fot_cost_pot(X, Y, Ts):
    """Feature cost of ``get_feature_coupling_pot`` (``MRI_PET_OT_nojax.py:121-136``).

    M_kl = sum_ij |X_ik - Y_jl|^2 Ts_ij, with w1 = Ts.sum(1), w2 = Ts.sum(0).
    No normalisation of Ts or of M.
    """
    # dtype promotion exactly as the reference: squares in the input dtype
    # (float32 embeddings), products promoted to float64 by the float64 Ts.
    X = np.asarray(X)
    Y = np.asarray(Y)
    Ts = np.asarray(Ts, dtype=np.float64)
    w1 = Ts.sum(axis=1)
    w2 = Ts.sum(axis=0)
    t1 = (X ** 2).T @ w1
    t2 = (Y ** 2).T @ w2
    t3 = -2 * X.T @ Ts @ Y
    return t1[:, None] + t2[None, :] + t3


def fot_cost_ott(X, Y, Ts):
    """Feature cost of ``fot_numpy`` (``perturbot/perturbot/match/fot.py:108-128``).

    Ts is normalised to unit mass first; the marginals are taken with the axes
    *swapped* relative to ``fot_cost_pot`` (w1 = Ts.sum(0), w2 = Ts.sum(1)),
    exactly as the reference does -- only consistent when n == n'.
    """
    X = np.asarray(X)
    Y = np.asarray(Y)
    Ts = np.asarray(Ts, dtype=np.float64)
    Ts = Ts / Ts.sum()
    w1 = Ts.sum(axis=0)
    w2 = Ts.sum(axis=1)
    constC, hC1, hC2 = init_matrix(X.T, Y.T, w1, w2)
    return constC - np.dot(hC1, Ts).dot(hC2.T), Ts


# --------------------------------------------------------------------------
# Sinkhorn, kernel domain (POT / in-tree mirror semantics)
# --------------------------------------------------------------------------
def sinkhorn_knopp(a, b, M=None, reg=None, K=None, numItermax=1000, stopThr=1e-9,
                   err_norm="l2", check_every=10, log=False, check_phase=1, u0=None, v0=None):
    """Kernel-domain Sinkhorn-Knopp.

    Follows ``perturbot/perturbot/match/utils.py:6-115`` line by line (init
    ``u=1/n, v=1/m`` :36-40; ``Kp`` :45; ``v=b/(K^T u)``, ``u=1/(Kp v)`` :51-53;
    zero/NaN/Inf guard restoring the previous duals :55-79; error every 10th
    iteration on the column marginal :80-89; return ``diag(u) K diag(v)``
    :111-115).

    ``err_norm="l2sq"`` + ``stop "<="`` is the in-tree mirror (squared norm,
    ``while err > stopThr``); ``err_norm="l2"`` + strict ``<`` is POT 0.9.6
    ``sinkhorn_knopp`` as called at ``MRI_PET_OT_nojax.py:143`` (upstream,
    unverifiable here).  Pass either ``K`` (mirror) or ``(M, reg)`` (POT).
    ``log["n_iter"]`` is the number of completed (v,u) updates; ``log["niter"]``
    is POT's loop index at exit.

    ``check_phase`` (default 1 = the reference: ``cpt % check_every == 0`` with the
    0-based counter) moves the check to 1-based iterations with
    ``it % check_every == check_phase % check_every``; ``check_phase=0`` with
    ``err_norm="l1"`` and ``u0 = v0 = 1`` is ott's rule evaluated in the kernel domain
    (the same map as ``sinkhorn_log`` while ``K`` has not underflowed -- used for
    problems too large for the log-domain oracle to finish in seconds).
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if K is None:
        M = np.asarray(M, dtype=np.float64)
        K = np.exp(M / (-reg))
    else:
        K = np.asarray(K, dtype=np.float64)
    n, m = len(a), len(b)
    u = np.ones(n) / n if u0 is None else np.asarray(u0, dtype=np.float64).copy()
    v = np.ones(m) / m if v0 is None else np.asarray(v0, dtype=np.float64).copy()
    Kp = (1.0 / a).reshape(-1, 1) * K
    errs = []
    err = 1.0
    cpt = 0
    flag = 0
    while cpt < numItermax:
        uprev, vprev = u, v
        KtU = K.T @ u
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            v = b / KtU
            u = 1.0 / (Kp @ v)
        if (np.any(KtU == 0) or np.any(np.isnan(u)) or np.any(np.isnan(v))
                or np.any(np.isinf(u)) or np.any(np.isinf(v))):
            u, v = uprev, vprev
            flag = 1  # "numerical errors": previous iterate returned
            break
        if (cpt + 1) % check_every == check_phase % check_every:
            # column marginal of diag(u) K diag(v), same op order as utils.py:88
            col = np.sum(u.reshape(-1, 1) * (K * v), axis=0)
            if err_norm == "l2sq":
                err = float(np.linalg.norm(col - b) ** 2)
            elif err_norm == "l2":
                err = float(np.linalg.norm(col - b))
            elif err_norm == "l1":
                err = float(np.abs(col - b).sum())
            else:
                raise ValueError(err_norm)
            errs.append(err)
            stop = err <= stopThr if err_norm == "l2sq" else err < stopThr
            if stop:
                cpt += 1
                break
        cpt += 1
    P = u.reshape(-1, 1) * K * v.reshape(1, -1)
    if log:
        return P, {"err": errs, "u": u, "v": v, "n_iter": cpt, "niter": max(cpt - 1, 0),
                   "numerical_error": flag}
    return P


# --------------------------------------------------------------------------
# Sinkhorn, log domain (the engine's own arithmetic, in float64)
# --------------------------------------------------------------------------
def _lse(x, axis):
    mx = np.max(x, axis=axis, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    return (np.log(np.sum(np.exp(x - mx), axis=axis, keepdims=True)) + mx).squeeze(axis)


def plan_from_potentials(C, f, g, eps):
    return np.exp((f[:, None] + g[None, :] - np.asarray(C, dtype=np.float64)) / eps)


def sinkhorn_log(C, a, b, eps, max_iter=1000, tol=1e-9, err_norm="l2", check_every=10,
                 check_phase=1, stop_inclusive=False, f0=None, g0=None, log=False, mask=None):
    """Log-domain Sinkhorn with a pluggable stopping rule.

    One iteration = ``g <- eps log b - eps LSE_i((f_i - C_ij)/eps)`` then
    ``f <- eps log a - eps LSE_j((g_j - C_ij)/eps)`` -- the same map as the
    kernel-domain ``v`` then ``u`` update with ``f = eps log u``, ``g = eps log v``
    (SURVEY.md Appendix A).  After iteration ``it`` (1-based) the column marginal
    error of ``exp((f+g-C)/eps)`` is evaluated when
    ``it % check_every == check_phase % check_every``:
    ``check_phase=1`` reproduces POT / the mirror (``cpt % 10 == 0`` with a
    0-based counter), ``check_phase=0`` reproduces ott's ``inner_iterations``.
    """
    C = np.asarray(C, dtype=np.float64)
    if mask is not None:  # plan restricted to the support of `mask`: no mass where it is 0
        C = np.where(np.asarray(mask) > 0, C, np.inf)
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n, m = C.shape
    f = np.zeros(n) if f0 is None else np.asarray(f0, dtype=np.float64).copy()
    g = np.zeros(m) if g0 is None else np.asarray(g0, dtype=np.float64).copy()
    la, lb = np.log(a), np.log(b)
    errs = []
    err = np.inf
    it = 0
    converged = False
    while it < max_iter:
        g = eps * lb - eps * _lse((f[:, None] - C) / eps, axis=0)
        f = eps * la - eps * _lse((g[None, :] - C) / eps, axis=1)
        it += 1
        if it % check_every == check_phase % check_every:
            col = np.exp((f[:, None] + g[None, :] - C) / eps).sum(axis=0)
            d = col - b
            if err_norm == "l2sq":
                err = float(d @ d)
            elif err_norm == "l2":
                err = float(np.sqrt(d @ d))
            elif err_norm == "l1":
                err = float(np.abs(d).sum())
            else:
                raise ValueError(err_norm)
            errs.append(err)
            if (err <= tol) if stop_inclusive else (err < tol):
                converged = True
                break
    with np.errstate(invalid="ignore"):
        P = plan_from_potentials(C, f, g, eps)
    if log:
        return P, {"err": errs, "f": f, "g": g, "n_iter": it, "converged": converged}
    return P


def rows_given_g(C_rows, a_rows, g, eps):
    """The f update and plan rows that follow from a given column potential g (float64): after any completed
    iteration f_i = eps log a_i - eps LSE_j((g_j - C_ij)/eps) and P_ij = exp((f_i + g_j - C_ij)/eps).  bench.py
    uses it to check rows of a 65536-column solve that no CPU can iterate itself: the converged g comes from the
    GPU, the row arithmetic (cost rows, logsumexp, plan entries) is redone here in float64."""
    C_rows = np.asarray(C_rows, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    f = eps * np.log(np.asarray(a_rows, dtype=np.float64)) - eps * _lse((g[None, :] - C_rows) / eps, axis=1)
    return f, np.exp((f[:, None] + g[None, :] - C_rows) / eps)


def sinkhorn_log_ott(M, eps, a=None, b=None, max_iterations=2000, threshold=1e-3,
                     inner_iterations=10, scale_cost="max_cost", log=False, mask=None):
    """ott-jax 0.6.0 ``linear.solve(Geometry(cost_matrix=M, epsilon=eps,
    scale_cost="max_cost"), max_iterations=N).matrix`` as called at
    ``perturbot/perturbot/match/fot.py:129-134`` (upstream, unverifiable here):
    cost divided by its max, eps absolute on the scaled cost, zero-initialised
    potentials, g-then-f updates, L1 error of the b-marginal every
    ``inner_iterations``, stop when ``err < threshold``.
    """
    M = np.asarray(M, dtype=np.float64)
    n, m = M.shape
    a = np.ones(n) / n if a is None else np.asarray(a, dtype=np.float64)
    b = np.ones(m) / m if b is None else np.asarray(b, dtype=np.float64)
    if scale_cost == "max_cost":
        C = M / M.max()
    elif scale_cost in (None, 1.0, "none"):
        C = M
    else:
        raise ValueError(scale_cost)
    P, lg = sinkhorn_log(C, a, b, eps, max_iter=max_iterations, tol=threshold, err_norm="l1",
                         check_every=inner_iterations, check_phase=0, log=True, mask=mask)
    lg["scaled_cost"] = C
    return (P, lg) if log else P


# --------------------------------------------------------------------------
# the reference's callers (a1, a3)
# --------------------------------------------------------------------------
def get_feature_coupling_pot(data, Ts, eps=5e-3, numItermax=2000, stopThr=1e-9, err_norm="l2"):
    """``MRI_PET_OT_nojax.py:91-145``: sorted-label concat, block-diagonal Ts,
    feature cost, uniform feature marginals, ``ot.sinkhorn(a,b,M,reg=eps,
    numItermax=2000)``.  Returns ``(Tv, {})``."""
    X_dict, Y_dict = data
    keys = sorted(X_dict.keys())
    X = np.concatenate([X_dict[l] for l in keys])
    Y = np.concatenate([Y_dict[l] for l in keys])
    if isinstance(Ts, dict):
        Ts = _ts_from_dict_sorted(X_dict, Y_dict, Ts)
    M = fot_cost_pot(X, Y, Ts)
    a = np.ones(X.shape[1]) / X.shape[1]
    b = np.ones(Y.shape[1]) / Y.shape[1]
    Tv = sinkhorn_knopp(a, b, M=M, reg=eps, numItermax=numItermax, stopThr=stopThr,
                        err_norm=err_norm)
    return Tv, {}


def fot_bcd_ott(X1, X2, Ts, reg2, niter=2000, log=False):
    """BCD shell of ``fot_numpy`` (``perturbot/perturbot/match/fot.py:104-152``):
    Ts fixed => the cost is identical every round; exit when
    ``|Tv - Tv_old|_F < 1e-16`` or ``|cost_old - cost| < 1e-7`` (:145)."""
    M, _ = fot_cost_ott(X1, X2, Ts)
    d1, d2 = X1.shape[1], X2.shape[1]
    Tv = np.ones((d1, d2)) / (d1 * d2)
    cost = np.inf
    costs = []
    rounds = 0
    for _ in range(niter):
        Tv_old, cost_old = Tv, cost
        Tv = sinkhorn_log_ott(M, reg2)
        delta = np.linalg.norm(Tv - Tv_old)
        cost = float(np.sum(M * Tv))
        costs.append(cost)
        rounds += 1
        if delta < 1e-16 or abs(cost_old - cost) < 1e-7:
            break
    if log:
        return Tv, cost, {"cost": costs, "rounds": rounds}
    return Tv, cost


def get_coupling_fot(data, Ts, eps=5e-3):
    """``perturbot/perturbot/match/fot.py:155-220`` (first-seen label order)."""
    X_dict, Y_dict = data
    keys = list(X_dict.keys())
    if isinstance(Ts, dict):
        Ts = mdict_to_matrix(
            Ts,
            np.concatenate([np.ones(X_dict[l].shape[0]) * l for l in keys]),
            np.concatenate([np.ones(Y_dict[l].shape[0]) * l for l in keys]),
        )
    X = np.concatenate([X_dict[l] for l in keys])
    Y = np.concatenate([Y_dict[l] for l in keys])
    Tv, cost, lg = fot_bcd_ott(X, Y, Ts, reg2=eps, niter=2000, log=True)
    return Tv, lg


# --------------------------------------------------------------------------
# label-constrained entropic COOT
# --------------------------------------------------------------------------
def cotl_sinkhorn(X_dict, Y_dict, reg=5e-3, niter=2000, log=False):
    """``cotl_numpy(algo="sinkhorn", algo2="sinkhorn")`` (``perturbot/perturbot/match/cot_labels.py:14-225``), the
    solver behind ``get_coupling_cotl_sinkhorn`` (:283-341).  Block coordinate descent: per label a sample coupling
    on ``M_k = constC_s - hC1_s Tv hC2_s^T`` (:172), then ONE feature coupling on the sum over labels of
    ``constC_v - hC1_v Ts_k hC2_v^T`` (:190-193), both through ott ``linear.solve(Geometry(cost_matrix, epsilon=reg,
    scale_cost="max_cost"), max_iterations=2000)`` -- the feature solve also uses ``reg``, ``reg2`` is never read
    (:201).  Default weights (:107-127): features weighted by their column sums when the data are non-negative,
    else uniform; samples uniform per label.  ``Tsold = Ts`` aliases the dict (:162), so ``delta`` only sees the
    change of ``Tv`` (:209-211); exit on ``delta < 1e-16`` or ``|cost_old - cost| < 1e-7`` (:219).  PINNED against the
    reference function executed in the build container (``tests/golden/cotl_sinkhorn.npz``)."""
    labels = list(X_dict.keys())
    X_dict = {k: np.asarray(X_dict[k], dtype=np.float64) for k in labels}
    Y_dict = {k: np.asarray(Y_dict[k], dtype=np.float64) for k in labels}
    X = np.concatenate([X_dict[k] for k in labels], axis=0)
    Y = np.concatenate([Y_dict[k] for k in labels], axis=0)
    v1 = X.sum(0) / X.sum() if (X >= 0).all() else np.ones(X.shape[1]) / X.shape[1]
    v2 = Y.sum(0) / Y.sum() if (Y >= 0).all() else np.ones(Y.shape[1]) / Y.shape[1]
    w1 = {k: np.ones(X_dict[k].shape[0]) / X_dict[k].shape[0] for k in labels}
    w2 = {k: np.ones(Y_dict[k].shape[0]) / Y_dict[k].shape[0] for k in labels}
    Ts = {k: np.outer(w1[k], w2[k]) for k in labels}
    d1, d2 = X.shape[1], Y.shape[1]
    Tv = np.ones((d1, d2)) / (d1 * d2)
    cs, cv = {}, {}
    for k in labels:
        cs[k] = init_matrix(X_dict[k], Y_dict[k], v1, v2)
        cv[k] = init_matrix(X_dict[k].T, Y_dict[k].T, w1[k], w2[k])
    cost = np.inf
    costs = []
    for _ in range(niter):
        Tv_old, cost_old = Tv, cost
        for k in labels:
            constC, hC1, hC2 = cs[k]
            Ts[k] = sinkhorn_log_ott(constC - hC1 @ Tv @ hC2.T, reg)
        M = 0
        for k in labels:
            constC, hC1, hC2 = cv[k]
            M = M + constC - hC1 @ Ts[k] @ hC2.T
        Tv = sinkhorn_log_ott(M, reg)
        if not abs(Tv.sum() - 1.0) < 1e-8:
            Tv = Tv / Tv.sum()
        delta = np.linalg.norm(Tv - Tv_old)
        cost = float(np.sum(M * Tv))
        costs.append(cost)
        if delta < 1e-16 or abs(cost_old - cost) < 1e-7:
            break
    if log:
        return Ts, Tv, cost, {"cost": costs}
    return Ts, Tv, cost


# --------------------------------------------------------------------------
# entropic Gromov-Wasserstein sample couplings
# --------------------------------------------------------------------------
def egw_ott(X, Y, eps=5e-3, gw_max_iterations=2000, sinkhorn_max_iterations=2000, gw_threshold=1e-3,
            gw_min_iterations=5, sk_threshold=1e-3, sk_check_every=10, mask=None):
    """ott-jax 0.6.0 ``GromovWasserstein(epsilon=eps, max_iterations=..., linear_solver=Sinkhorn(max_iterations=...))``
    on ``QuadraticProblem(PointCloud(x, x, scale_cost="max_cost"), PointCloud(y, y, scale_cost="max_cost"))``
    (``MRI_PET_OT_OT_per_epoch_attn.py:155-175``).  PARITY UNPINNED: ott is not in the tree; this restates its
    published algorithm -- squared-Euclidean geometries divided by their maximum, square loss decomposition
    ``f1(x) = x^2, f2(y) = y^2, h1(x) = x, h2(y) = 2y``, uniform marginals, ``T0 = a b^T``; every outer iteration
    linearises at the current coupling (marginal terms from the coupling's own marginals), solves the linear
    problem with the log-domain Sinkhorn of ``sinkhorn_log_ott`` (absolute epsilon, L1 error of the b-marginal
    every 10 iterations, threshold 1e-3) warm-started from the previous potentials, and records
    ``cost = <a, f> + <b, g>``; it stops once ``iteration >= min_iterations`` and
    ``isclose(costs[i - 2], costs[i - 1], rtol=threshold)``.  Returns ``(T, log)``.

    ``mask`` (n x m of 0/1, optional): the label-aware form called at ``perturbot/perturbot/match/ott_egwl.py:77-105``
    (``QuadraticProblem(..., labels_a, labels_b, n_labels, block_diag_mat)`` of a modified OTT that is NOT in the
    tree, version unknown).  Restated from that function's docstring (``T_ij > 0 => l_{x_i} = l_{y_j}``, :37): the
    coupling is restricted to the support of ``mask`` -- start coupling ``(a b^T) * mask``, every linearised cost
    infinite outside it -- with the global uniform marginals kept.  PARITY UNPINNED."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n, m = X.shape[0], Y.shape[0]
    C1 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    C2 = ((Y[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
    C1 = C1 / C1.max() if C1.max() > 0 else C1
    C2 = C2 / C2.max() if C2.max() > 0 else C2
    a = np.full(n, 1.0 / n)
    b = np.full(m, 1.0 / m)
    la, lb = np.log(a), np.log(b)
    T = np.outer(a, b)
    if mask is not None:
        mask = np.asarray(mask) > 0
        T = T * mask
    f = np.zeros(n)
    g = np.zeros(m)
    costs = []
    inner_total = 0
    inner_conv = False
    outer_conv = False
    while len(costs) < gw_max_iterations:
        M = ((C1 * C1) @ T.sum(1))[:, None] + ((C2 * C2) @ T.sum(0))[None, :] - 2.0 * C1 @ T @ C2
        if mask is not None:
            M = np.where(mask, M, np.inf)
        it = 0
        inner_conv = False
        while it < sinkhorn_max_iterations:
            g = eps * lb - eps * _lse((f[:, None] - M) / eps, axis=0)
            f = eps * la - eps * _lse((g[None, :] - M) / eps, axis=1)
            it += 1
            if it % sk_check_every == 0:
                err = np.abs(np.exp((f[:, None] + g[None, :] - M) / eps).sum(0) - b).sum()
                if err < sk_threshold:
                    inner_conv = True
                    break
        inner_total += it
        T = np.exp((f[:, None] + g[None, :] - M) / eps)
        costs.append(float(a @ f + b @ g))
        if not np.isfinite(costs[-1]):
            break
        k = len(costs)
        if k >= 2:
            outer_conv = bool(abs(costs[-2] - costs[-1]) <= 1e-8 + gw_threshold * abs(costs[-1]))
            if outer_conv and k >= gw_min_iterations:
                break
    return T, {"n_iters_outer": len(costs), "converged_outer": outer_conv, "converged_inner": inner_conv,
               "inner_iterations": inner_total, "GW cost": costs[-1], "costs": costs}


def get_coupling_egw_ott_fixed(data, eps=5e-3, gw_max_iterations=2000, sinkhorn_max_iterations=2000):
    """``MRI_PET_OT_OT_per_epoch_attn.py:129-186``: one coupling per label, NaN features mapped to 0 (:148-151)."""
    X_dict, Y_dict = data
    Ts, log = {}, {}
    for l in X_dict.keys():
        T, lg = egw_ott(np.nan_to_num(np.asarray(X_dict[l], dtype=np.float64)),
                        np.nan_to_num(np.asarray(Y_dict[l], dtype=np.float64)), eps, gw_max_iterations,
                        sinkhorn_max_iterations)
        Ts[l] = T
        log[l] = lg
    return Ts, log


def _concat_with_labels(X_dict, Y_dict):
    keys = list(X_dict.keys())
    Xs = np.concatenate([np.asarray(X_dict[l], dtype=np.float64) for l in keys], axis=0)
    Xt = np.concatenate([np.asarray(Y_dict[l], dtype=np.float64) for l in keys], axis=0)
    sl = np.concatenate([np.repeat(l, np.asarray(X_dict[l]).shape[0]) for l in keys])
    tl = np.concatenate([np.repeat(l, np.asarray(Y_dict[l]).shape[0]) for l in keys])
    return Xs, Xt, sl, tl


def block_diag_mask(labels_a, labels_b):
    """``create_block_diag_mat`` (``perturbot/perturbot/match/ott_egwl.py:16-22``): 1 where the labels agree."""
    labels_a, labels_b = np.asarray(labels_a), np.asarray(labels_b)
    out = np.zeros((len(labels_a), len(labels_b)))
    for l in np.unique(labels_a):
        out[np.ix_(np.where(labels_a == l)[0], np.where(labels_b == l)[0])] = 1.0
    return out


def get_coupling_egw_all_ott(data, eps=5e-3):
    """``perturbot/perturbot/match/ott_egwl.py:209-297``: ONE entropic GW problem over all samples (labels
    disregarded), ``GromovWasserstein(epsilon=eps, max_iterations=1000)`` with the default inner ``Sinkhorn``."""
    Xs, Xt, _, _ = _concat_with_labels(*data)
    T, lg = egw_ott(Xs, Xt, eps, gw_max_iterations=1000, sinkhorn_max_iterations=2000)
    return T, lg


def get_coupling_egw_labels_ott(data, eps=5e-3):
    """``perturbot/perturbot/match/ott_egwl.py:25-127``: the label-constrained entropic GW (block-diagonal support,
    see ``egw_ott(mask=...)``), ``max_iterations=2000`` outer and inner; returns the per-label diagonal blocks."""
    Xs, Xt, sl, tl = _concat_with_labels(*data)
    T, lg = egw_ott(Xs, Xt, eps, gw_max_iterations=2000, sinkhorn_max_iterations=2000,
                    mask=block_diag_mask(sl, tl))
    return {l: T[sl == l, :][:, tl == l] for l in np.unique(sl)}, lg


def get_coupling_leot_ott(data, eps=5e-3):
    """``perturbot/perturbot/match/ott_egwl.py:375-454``: label-constrained entropic OT.  Squared-Euclidean cost over
    all samples divided by its maximum (``PointCloud(scale_cost="max_cost").cost_matrix``, :426-427), then ott
    ``Sinkhorn()`` on ``LinearProblem(geom, labels_a, labels_b)`` of the modified OTT (not in the tree): restated as
    the plan restricted to matching labels with the global uniform marginals kept.  PARITY UNPINNED."""
    Xs, Xt, sl, tl = _concat_with_labels(*data)
    C = sqeuclid_cost(Xs, Xt)
    C = C / C.max()
    T, lg = sinkhorn_log_ott(C, eps, scale_cost=None, mask=block_diag_mask(sl, tl), log=True)
    return {l: T[sl == l, :][:, tl == l] for l in np.unique(sl)}, lg


# --------------------------------------------------------------------------
# epilogues (a6)
# --------------------------------------------------------------------------
def plan_guard_rownorm(T):
    """NaN -> 1e-8, then row-normalise with zero-row guard
    (``MRI_PET_OT_nojax.py:704-715``)."""
    T = np.asarray(T, dtype=np.float64).copy()
    T[np.isnan(T)] = 1e-8
    rs = T.sum(axis=1, keepdims=True)
    rs[rs == 0] = 1e-8
    return T / rs


def apply_plan_T(V, T):
    """``pet @ T.t()`` (``MRI_PET_OT_OT_per_epoch_attn.py:728``,
    ``MRI_PET_OT_nojax.py:718``)."""
    return np.asarray(V, dtype=np.float64) @ np.asarray(T, dtype=np.float64).T


def barycentric(P, Y):
    """``(T / rowsum) @ Y`` with ``rowsum == 0 -> 1e-30``
    (``perturbot/perturbot/eval/match.py:202-206``)."""
    P = np.asarray(P, dtype=np.float64)
    marg = P.sum(axis=-1)
    marg = np.where(marg == 0, 1e-30, marg)
    return (P / marg[:, None]) @ np.asarray(Y, dtype=np.float64)


def cosine_loss(x, y):
    """``1 - mean_i cos(x_i, y_i)`` with F.normalize's 1e-12 floor
    (``MRI_PET_OT_nojax.py:552-560``)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    yn = y / np.maximum(np.linalg.norm(y, axis=1, keepdims=True), 1e-12)
    num = (xn * yn).sum(1)
    den = np.maximum(np.linalg.norm(xn, axis=1) * np.linalg.norm(yn, axis=1), 1e-8)
    return float(1.0 - (num / den).mean())


def ot_cost(P, C):
    """Transport cost <P, C> (``perturbot/perturbot/match/fot.py:137``)."""
    return float(np.sum(np.asarray(P, dtype=np.float64) * np.asarray(C, dtype=np.float64)))


def envelope_grads(X, Y, P):
    """Envelope-theorem gradients of <P, C(X,Y)> for the squared-Euclidean cost
    with P held fixed: dX = 2 (diag(P1) X - P Y), dY = 2 (diag(P^T1) Y - P^T X).
    New capability (the reference never differentiates through OT,
    ``MRI_PET_OT_nojax.py:683-684``); validated against torch-fp64 autograd in
    tests, not against the reference."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    dX = 2.0 * (P.sum(1)[:, None] * X - P @ Y)
    dY = 2.0 * (P.sum(0)[:, None] * Y - P.T @ X)
    return dX, dY


def foscttm(Y_pred, Y_true):
    """Fraction of samples closer than the true match (``perturbot/perturbot/eval/utils.py:18-45``): per row,
    Euclidean distances to every true sample, sorted; rank = mean position of the true match's distance."""
    Y_pred = np.asarray(Y_pred, dtype=np.float64)
    Y_true = np.asarray(Y_true, dtype=np.float64)
    n = Y_pred.shape[0]
    fracs = []
    for i in range(n):
        dist = np.sqrt(np.sum(np.square(Y_pred[i, :] - Y_true), axis=1))
        rank = np.where(np.sort(dist) == dist[i])[0].mean()
        fracs.append(float(rank) / (n - 1))
    return fracs


def group_features_by_label(y, p, max_samples_per_label=None):
    """``MRI_PET_OT_OT_per_epoch_attn.py:918-937``: rows of p bucketed by label (np.unique order), truncated."""
    y = np.asarray(y)
    p = np.asarray(p)
    out = {}
    for label in np.unique(y):
        arr = p[y == label]
        if max_samples_per_label is not None and max_samples_per_label > 0 and arr.shape[0] > max_samples_per_label:
            arr = arr[:max_samples_per_label]
        out[int(label)] = arr
    return out
