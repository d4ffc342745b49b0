"""The oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import ot_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_knopp_matches_reference_mirror_c1(golden_dir):
    g = _load(golden_dir, "c1_sample_64.npz")
    a = b = np.ones(64) / 64
    K = np.exp(-g["C"] / float(g["eps"]))
    P, lg = orc.sinkhorn_knopp(a, b, K=K, numItermax=200, stopThr=0.0, err_norm="l2sq", log=True)
    # same BLAS calls in the same order as perturbot/match/utils.py:51-53 -> bit-exact
    assert np.array_equal(P, g["P200"])
    assert np.array_equal(np.array(lg["err"]), g["err200"])
    assert lg["n_iter"] == 200
    P, lg = orc.sinkhorn_knopp(a, b, K=K, numItermax=2000, stopThr=1e-9, err_norm="l2sq", log=True)
    assert np.array_equal(P, g["Pconv"])
    assert len(lg["err"]) == len(g["errconv"]) and lg["n_iter"] == 11


def test_cost_matches_golden(golden_dir):
    g = _load(golden_dir, "c1_sample_64.npz")
    X, Y = orc.synthetic_embeddings(64, 64, 512, config_index=0)
    assert np.array_equal(X, g["X"]) and np.array_equal(Y, g["Y"])
    assert np.array_equal(orc.sqeuclid_cost(X, Y), g["C"])


def test_log_domain_equals_kernel_domain(golden_dir):
    g = _load(golden_dir, "c1_sample_64.npz")
    a = b = np.ones(64) / 64
    eps = float(g["eps"])
    f0 = eps * np.log(np.ones(64) / 64)  # u0 = 1/n  (utils.py:36-40)
    P, lg = orc.sinkhorn_log(g["C"], a, b, eps, max_iter=200, tol=0.0, err_norm="l2sq",
                             check_phase=1, stop_inclusive=False, f0=f0, log=True)
    np.testing.assert_allclose(P, g["P200"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(eps * np.log(g["u200"]), lg["f"], rtol=0, atol=1e-12)
    P, lg = orc.sinkhorn_log(g["C"], a, b, eps, max_iter=2000, tol=1e-9, err_norm="l2sq",
                             check_phase=1, stop_inclusive=True, f0=f0, log=True)
    assert lg["n_iter"] == 11 and lg["converged"]
    np.testing.assert_allclose(P, g["Pconv"], rtol=1e-10, atol=0)


def test_feature_coupling_pot_512(golden_dir):
    g = _load(golden_dir, "fot_pot_512.npz")
    Tv, _ = orc.get_feature_coupling_pot(({0: g["X"]}, {0: g["Y"]}), {0: np.eye(64) / 64},
                                         eps=float(g["eps"]), err_norm="l2sq")
    np.testing.assert_allclose(Tv, g["Tv"], rtol=1e-12, atol=0)


def test_feature_coupling_pot_labels(golden_dir):
    g = _load(golden_dir, "fot_pot_labels.npz")
    Xd = {1: g["X1"], 0: g["X0"]}
    Yd = {1: g["Y1"], 0: g["Y0"]}
    Ts = {1: g["Ts1"], 0: g["Ts0"]}
    Tv, _ = orc.get_feature_coupling_pot((Xd, Yd), Ts, eps=float(g["eps"]), err_norm="l2sq")
    np.testing.assert_allclose(Tv, g["Tv"], rtol=1e-12, atol=0)


def test_fot_ott_512(golden_dir):
    g = _load(golden_dir, "fot_ott_512.npz")
    Tv, lg = orc.get_coupling_fot(({0: g["X"]}, {0: g["Y"]}), {0: np.eye(64) / 64},
                                  eps=float(g["eps"]))
    np.testing.assert_allclose(Tv, g["Tv"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(lg["cost"], g["costs"], rtol=1e-12)
    assert lg["rounds"] == 2  # Ts fixed => two identical solves (fot.py:145)


def test_helpers(golden_dir):
    g = _load(golden_dir, "helpers.npz")
    constC, hC1, hC2 = orc.init_matrix(g["X1"], g["X2"], g["v1"], g["v2"])
    np.testing.assert_allclose(constC, g["constC"], rtol=1e-14)
    assert np.array_equal(hC1, g["hC1"]) and np.array_equal(hC2, g["hC2"])
    Mtot = orc.mdict_to_matrix({0: g["M0"], 1: g["M1"], 2: g["M2"]}, g["src"], g["tgt"])
    assert np.array_equal(Mtot, g["Mtot"])


def test_feature_coupling_pot_2048_summary(golden_dir):
    g = _load(golden_dir, "fot_pot_2048_summary.npz")
    Xb, Yb = orc.synthetic_embeddings(128, 128, 2048, config_index=10)
    Xb, Yb = np.abs(Xb) * float(g["scale"]), np.abs(Yb) * float(g["scale"])
    Tv, _ = orc.get_feature_coupling_pot(({0: Xb}, {0: Yb}), {0: np.eye(128) / 128},
                                         eps=float(g["eps"]), err_norm="l2sq")
    np.testing.assert_allclose(Tv[::256], g["rows"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(Tv.sum(1), g["rowsum"], rtol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(Tv), g["fro"], rtol=1e-12)


@pytest.mark.parametrize("norm,phase,incl", [("l2sq", 1, True), ("l2", 1, False), ("l1", 0, False)])
def test_stopping_rules_consistent(norm, phase, incl):
    """kernel- and log-domain solvers stop at the same iteration under each rule."""
    rng = np.random.default_rng(3)
    centers = rng.standard_normal((4, 8))
    X = centers[rng.integers(0, 4, 96)] + 0.3 * rng.standard_normal((96, 8))
    Y = centers[rng.integers(0, 4, 80)] + 0.3 * rng.standard_normal((80, 8))
    C = orc.sqeuclid_cost(X, Y)
    C /= C.max()
    a = np.ones(96) / 96
    b = np.ones(80) / 80
    eps = 0.02
    tol = {"l2sq": 1e-9, "l2": 1e-7, "l1": 1e-3}[norm]
    f0 = eps * np.log(a)
    Pl, ll = orc.sinkhorn_log(C, a, b, eps, max_iter=2000, tol=tol, err_norm=norm,
                              check_phase=phase, stop_inclusive=incl, f0=f0, log=True)
    if phase == 1:
        Pk, lk = orc.sinkhorn_knopp(a, b, M=C, reg=eps, numItermax=2000, stopThr=tol,
                                    err_norm=norm, log=True)
        assert lk["n_iter"] == ll["n_iter"] and ll["n_iter"] > 20
        np.testing.assert_allclose(Pl, Pk, rtol=1e-8, atol=1e-300)
    assert ll["converged"]


def test_epilogues():
    rng = np.random.default_rng(0)
    T = rng.random((6, 6))
    T[2] = 0.0
    T[3, 1] = np.nan
    Tn = orc.plan_guard_rownorm(T)
    assert np.isfinite(Tn).all()
    np.testing.assert_allclose(np.delete(Tn.sum(1), 2), 1.0)
    V = rng.standard_normal((4, 6))
    np.testing.assert_allclose(orc.apply_plan_T(V, Tn), V @ Tn.T)
    P = rng.random((5, 7))
    P[1] = 0
    Yv = rng.standard_normal((7, 3))
    B = orc.barycentric(P, Yv)
    assert np.allclose(B[1], 0) and np.allclose(B[0], P[0] @ Yv / P[0].sum())
    x = rng.standard_normal((4, 6))
    assert abs(orc.cosine_loss(x, x)) < 1e-12
    assert abs(orc.cosine_loss(x, -x) - 2) < 1e-12


def test_envelope_grads_match_autograd():
    import torch
    rng = np.random.default_rng(1)
    X = rng.standard_normal((7, 5))
    Y = rng.standard_normal((9, 5))
    P = rng.random((7, 9))
    Xt = torch.tensor(X, requires_grad=True)
    Yt = torch.tensor(Y, requires_grad=True)
    C = (Xt * Xt).sum(1)[:, None] + (Yt * Yt).sum(1)[None, :] - 2 * Xt @ Yt.T
    (torch.tensor(P) * C).sum().backward()
    dX, dY = orc.envelope_grads(X, Y, P)
    np.testing.assert_allclose(dX, Xt.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dY, Yt.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_foscttm_and_grouping_match_reference(golden_dir):
    """oracle.foscttm / group_features_by_label against outputs of the reference's own functions."""
    g = _load(golden_dir, "metrics_helpers.npz")
    np.testing.assert_allclose(orc.foscttm(g["pred"], g["true"]), g["foscttm"], rtol=0, atol=1e-15)
    grouped = orc.group_features_by_label(g["labels"], g["feats"], max_samples_per_label=20)
    assert sorted(grouped.keys()) == list(g["keys"])
    for k in grouped:
        assert np.array_equal(grouped[k], g[f"group{k}"])


def test_egw_oracle_recovers_an_isometric_copy_and_keeps_marginals():
    """egw_ott (restated ott GromovWasserstein, parity unpinned): structural checks only -- the coupling of a point
    cloud with a rotated, permuted, slightly noisy copy of itself is the permutation; marginals are uniform; the
    outer loop runs at least min_iterations and stops on the cost criterion."""
    rng = np.random.default_rng(0)
    n = 40
    X = rng.standard_normal((n, 16))
    Q, _ = np.linalg.qr(rng.standard_normal((16, 16)))
    perm = rng.permutation(n)
    Y = (X @ Q)[perm] + 0.01 * rng.standard_normal((n, 16))
    T, lg = orc.egw_ott(X, Y, eps=5e-3)
    assert lg["n_iters_outer"] >= 5 and lg["converged_outer"] and lg["converged_inner"]
    np.testing.assert_allclose(T.sum(1), np.full(n, 1.0 / n), rtol=1e-9)  # the f update is exact
    assert np.abs(T.sum(0) - 1.0 / n).sum() < 1e-3  # the inner stopping rule
    assert (T.argmax(1) == np.argsort(perm)).all()
    c = lg["costs"]
    assert abs(c[-2] - c[-1]) <= 1e-8 + 1e-3 * abs(c[-1])
    Ts, log = orc.get_coupling_egw_ott_fixed(({0: X[:10], 1: X[10:25]}, {0: Y[:12], 1: Y[12:30]}))
    assert Ts[0].shape == (10, 12) and Ts[1].shape == (15, 18) and set(log) == {0, 1}


def test_cotl_sinkhorn_matches_the_reference_function(golden_dir):
    """BCD shell of cotl_numpy (cot_labels.py:14-225) against the golden produced by the reference's own function:
    same number of rounds, same cost trace, same couplings."""
    g = np.load(os.path.join(golden_dir, "cotl_sinkhorn.npz"))
    keys = [int(k) for k in g["keys"]]
    Xd = {k: g[f"X{k}"] for k in keys}
    Yd = {k: g[f"Y{k}"] for k in keys}
    Ts, Tv, cost, lg = orc.cotl_sinkhorn(Xd, Yd, reg=float(g["reg"]), niter=2000, log=True)
    assert len(lg["cost"]) == len(g["costs"])
    np.testing.assert_allclose(lg["cost"], g["costs"], rtol=1e-10)
    np.testing.assert_allclose(Tv, g["Tv"], rtol=0, atol=1e-12)
    for k in keys:
        np.testing.assert_allclose(Ts[k], g[f"Ts{k}"], rtol=0, atol=1e-12)


def test_label_aware_ott_shells_match_reference_code(golden_dir):
    """get_coupling_egw_labels_ott / get_coupling_egw_all_ott / get_coupling_leot_ott
    (perturbot/perturbot/match/ott_egwl.py:25-127,209-297,375-454): the oracle's restatements against the golden
    produced by the reference's own function bodies (tests/golden/make_golden.py --ott-labels: concatenation order,
    label arrays, block-diagonal matrix, solver parameters and per-label slicing are the reference's code)."""
    g = np.load(os.path.join(golden_dir, "ott_labels.npz"))
    keys = [int(k) for k in g["keys"]]
    Xd = {k: g[f"X{k}"] for k in keys}
    Yd = {k: g[f"Y{k}"] for k in keys}
    eps = float(g["eps"])
    Tl, lgl = orc.get_coupling_egw_labels_ott((Xd, Yd), eps)
    Ta, lga = orc.get_coupling_egw_all_ott((Xd, Yd), eps)
    To, lgo = orc.get_coupling_leot_ott((Xd, Yd), eps)
    assert [int(k) for k in Tl.keys()] == sorted(keys) == [int(k) for k in To.keys()]  # np.unique order (:124)
    for k in keys:
        np.testing.assert_allclose(Tl[k], g[f"egwl_T{k}"], rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(To[k], g[f"leot_T{k}"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(Ta, g["egwa_T"], rtol=1e-10, atol=1e-300)
    assert lgl["n_iters_outer"] == int(g["egwl_log"][0]) and lga["n_iters_outer"] == int(g["egwa_log"][0])
    assert lgo["n_iter"] == int(g["leot_log"][0])
    # the constraint: all the mass sits on pairs of equal label, total mass 1
    assert abs(sum(v.sum() for v in Tl.values()) - 1.0) < 1e-3 and abs(sum(v.sum() for v in To.values()) - 1.0) < 1e-3
