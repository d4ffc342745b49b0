"""GPU parity of the round-2 additions to the reference's call surface (VERDICT r01 "missing" 1, 3, 4 and the
ADVICE item on fot_numpy): the label-aware / all-to-all ott call sites of perturbot/perturbot/match/ott_egwl.py, the
per-step plan guard of MRI_PET_OT_nojax.py:704-715, fot_numpy with the reference's own positional signature, and
the opt-in warm starts."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from oracle import ot_oracle as orc

import _parity

pytestmark = pytest.mark.gpu
RTOL = _parity.RTOL


def _dev(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=torch.float32)


def _golden_labels(golden_dir):
    g = np.load(os.path.join(golden_dir, "ott_labels.npz"))
    keys = [int(k) for k in g["keys"]]
    return g, keys, {k: g[f"X{k}"] for k in keys}, {k: g[f"Y{k}"] for k in keys}


def test_label_aware_ott_call_sites_match_reference_shell_golden(cuda_dev, golden_dir):
    """Goldens from the reference's own function bodies (inner solves = the oracle's ott restatement, unpinned)."""
    import b200ot
    g, keys, Xd, Yd = _golden_labels(golden_dir)
    eps = float(g["eps"])
    To, lgo = b200ot.get_coupling_leot_ott((Xd, Yd), eps)
    assert [int(k) for k in To.keys()] == sorted(keys)
    assert lgo["n_iters_outer"] == int(g["leot_log"][0]) and lgo["converged"] == bool(g["leot_log"][1])
    assert abs(lgo["OT cost"] - float(g["leot_log"][2])) < 1e-4 * max(1.0, abs(float(g["leot_log"][2])))
    for k in keys:
        assert isinstance(To[k], np.ndarray) and To[k].shape == g[f"leot_T{k}"].shape
        assert _parity.rel(To[k], g[f"leot_T{k}"], what=f"leot_{k}") < RTOL
    Tl, lgl = b200ot.get_coupling_egw_labels_ott((Xd, Yd), eps)
    assert [int(k) for k in Tl.keys()] == sorted(keys)
    assert lgl["n_iters_outer"] == int(g["egwl_log"][0]) and lgl["converged_outer"] == bool(g["egwl_log"][2])
    assert abs(lgl["GW cost"] - float(g["egwl_log"][3])) < 1e-4 * max(1.0, abs(float(g["egwl_log"][3])))
    for k in keys:
        assert _parity.rel(Tl[k], g[f"egwl_T{k}"], what=f"egwl_{k}") < RTOL
    Ta, lga = b200ot.get_coupling_egw_all_ott((Xd, Yd), eps)
    assert Ta.shape == g["egwa_T"].shape and lga["n_iters_outer"] == int(g["egwa_log"][0])
    assert _parity.rel(Ta, g["egwa_T"], what="egw_all") < RTOL
    # CUDA tensors in -> the couplings stay on the device
    Xc = {k: _dev(v, cuda_dev) for k, v in Xd.items()}
    Yc = {k: _dev(v, cuda_dev) for k, v in Yd.items()}
    Tc, _ = b200ot.get_coupling_leot_ott((Xc, Yc), eps)
    assert all(t.is_cuda for t in Tc.values())


def test_entropic_gw_labels_beyond_the_shared_memory_cap(cuda_dev):
    """get_coupling_egw_ott_fixed with a label of 100 x 90 samples (more than the 64 the one-CTA kernel holds):
    served by the dense path, same couplings and outer iteration count as the float64 restatement."""
    import b200ot
    rng = np.random.default_rng(17)
    X = rng.standard_normal((100, 12)).astype(np.float32)
    Q, _ = np.linalg.qr(rng.standard_normal((12, 12)))
    Y = ((X @ Q)[rng.permutation(100)][:90] + 0.02 * rng.standard_normal((90, 12))).astype(np.float32)
    small = (rng.standard_normal((20, 12)).astype(np.float32), rng.standard_normal((25, 12)).astype(np.float32))
    Ts, log = b200ot.get_coupling_egw_ott_fixed(({4: X, 1: small[0]}, {4: Y, 1: small[1]}), eps=5e-2)
    assert list(Ts.keys()) == [4, 1]
    Tr, lr = orc.egw_ott(X, Y, eps=5e-2)
    assert log[4]["n_iters_outer"] == lr["n_iters_outer"] and log[4]["converged_outer"] == lr["converged_outer"]
    assert _parity.rel(Ts[4], Tr, what="egw_dense_100x90") < RTOL
    Tr1, _ = orc.egw_ott(small[0], small[1], eps=5e-2)
    assert _parity.rel(Ts[1], Tr1, what="egw_smem_20x25") < RTOL


def test_per_step_plan_guard_rownorm(cuda_dev):
    """MRI_PET_OT_nojax.py:679-715 on device against the composition of the oracle's restatements
    (get_feature_coupling_pot pinned to the reference function, plan_guard_rownorm :704-715), plus the guard on a
    materialised plan with NaNs and an all-zero row."""
    import b200ot
    from b200ot import ops
    rng = np.random.default_rng(23)
    B, d = 8, 96
    mri = np.abs(rng.standard_normal((B, d))).astype(np.float32)
    pet = np.abs(rng.standard_normal((B, d)) + 0.2).astype(np.float32)
    Tv, _ = orc.get_feature_coupling_pot(({0: mri}, {0: pet}), {0: np.eye(B) / B}, eps=1e-2, stopThr=1e-6)
    ref = orc.plan_guard_rownorm(Tv)
    T = b200ot.per_step_feature_plan(_dev(mri, cuda_dev), _dev(pet, cuda_dev), eps=1e-2, stopThr=1e-6)
    assert T.is_cuda and T.shape == (d, d)
    assert _parity.rel(T.double().cpu().numpy(), ref, what="per_step_plan") < RTOL
    np.testing.assert_allclose(T.sum(1).cpu().numpy(), 1.0, rtol=1e-5)
    # the reference's forward applies it as pet @ T.t() (:718)
    Z = (_dev(pet, cuda_dev) @ T.t()).double().cpu().numpy()
    np.testing.assert_allclose(Z, orc.apply_plan_T(pet, ref), rtol=0, atol=RTOL * np.abs(orc.apply_plan_T(pet, ref)).max())
    # dense plan with NaNs and a zero row
    P = np.abs(rng.standard_normal((37, 53))).astype(np.float32)
    P[3, 7] = np.nan
    P[11, :] = 0.0
    G = ops.plan_guard_rownorm(T=_dev(P, cuda_dev)).double().cpu().numpy()
    np.testing.assert_allclose(G, orc.plan_guard_rownorm(P.astype(np.float64)), rtol=2e-6, atol=1e-12)
    assert np.all(G[11] == 0.0) and np.isfinite(G).all()


def test_fot_numpy_keeps_the_reference_signature(cuda_dev, golden_dir):
    """fot_numpy(X1, X2, Ts, v1, v2, niter, algo, reg, algo2, reg2, verbose, log, ...) (fot.py:14-28): positional
    order, defaults, printed lines, the cost returned without log=True."""
    import b200ot
    g = np.load(os.path.join(golden_dir, "fot_ott_512.npz"))
    Ts = np.eye(64) / 64
    buf = io.StringIO()
    with redirect_stdout(buf):
        Tv, cost = b200ot.fot_numpy(g["X"], g["Y"], Ts, None, None, 10, "sinkhorn", float(g["eps"]), "sinkhorn",
                                    float(g["eps"]))
    assert "Delta" in buf.getvalue() and "converged at iter" in buf.getvalue()  # verbose=True is the default
    assert _parity.rel(Tv, g["Tv"], what="fot_numpy_positional") < RTOL
    assert abs(cost - float(g["cost"])) < RTOL * abs(float(g["cost"]))
    Tv2, cost2, lg = b200ot.fot_numpy(g["X"], g["Y"], Ts, reg2=float(g["eps"]), niter=2000, log=True, verbose=False,
                                      algo="sinkhorn", algo2="sinkhorn", v1=np.ones(512) / 512, C_lin=None)
    assert lg["cost"][-1] == cost2 and len(lg["cost"]) == len(g["costs"]) and _parity.rel(Tv2, g["Tv"]) < RTOL
    with pytest.raises(b200ot.B200OTError):
        b200ot.fot_numpy(g["X"], g["Y"], Ts, verbose=False)              # reg2 = 0 (the reference default)
    with pytest.raises(b200ot.B200OTError):
        b200ot.fot_numpy(g["X"], g["Y"], Ts, reg2=0.01, random_init=True, verbose=False)
    # warm start from the previous solve's potentials: same coupling, fewer iterations
    Tv3, _, lg3 = b200ot.fot_numpy(g["X"], g["Y"], Ts, reg2=float(g["eps"]), log=True, verbose=False,
                                   warm_start=lg["potentials"])
    # both solves stop on the same rule (L1 < 1e-3): the warm-started iterate is a different point inside that band
    assert lg3["n_iters"] <= lg["n_iters"] and _parity.rel(Tv3, g["Tv"], elem_rtol=None) < 1e-2


def test_cotl_warm_start_saves_inner_iterations(cuda_dev, golden_dir):
    import b200ot
    g = np.load(os.path.join(golden_dir, "cotl_sinkhorn.npz"))
    keys = [int(k) for k in g["keys"]]
    Xd = {k: g[f"X{k}"] for k in keys}
    Yd = {k: g[f"Y{k}"] for k in keys}
    kw = dict(niter=30, algo="sinkhorn", reg=float(g["reg"]), algo2="sinkhorn", reg2=float(g["reg"]), verbose=False,
              log=True)
    _, Tv0, c0, l0 = b200ot.cotl_numpy(Xd, Yd, **kw)
    _, Tv1, c1, l1 = b200ot.cotl_numpy(Xd, Yd, warm_start=True, **kw)
    # checks come every 10 iterations, so a solve cannot finish in fewer than 10: the saving is bounded by that
    assert l1["inner_iterations"] <= l0["inner_iterations"], (l1["inner_iterations"], l0["inner_iterations"])
    assert abs(c1 - c0) < 5e-3 * abs(c0), (c1, c0)
    assert _parity.rel(Tv1, Tv0, elem_rtol=None) < 5e-2


def test_ot_cost_is_bit_reproducible(cuda_dev):
    from b200ot import ops
    rng = np.random.default_rng(4)
    C = _dev(rng.random((1500, 2300)), cuda_dev)
    f = _dev(rng.standard_normal(1500) * 0.05, cuda_dev)
    g = _dev(rng.standard_normal(2300) * 0.05, cuda_dev)
    vals = {float(ops.ot_cost(C, f, g, 0.3).item()) for _ in range(5)}
    assert len(vals) == 1
    P = orc.plan_from_potentials(C.double().cpu().numpy(), f.double().cpu().numpy(), g.double().cpu().numpy(), 0.3)
    ref = orc.ot_cost(P, C.double().cpu().numpy())
    assert abs(vals.pop() - ref) < 1e-5 * ref
