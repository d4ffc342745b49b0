"""GPU parity: the CUDA path (through the C ABI) against the float64 oracle and the golden
vectors produced by the reference's own code.  Run with ``-m gpu`` on a B200.

Tolerances (north star): plan / loss / fused embeddings within 1e-4 relative in fp32;
iteration counts exactly equal.  "Relative" for the plan is measured against the largest
entry of the reference plan (entries far below it carry no transport mass).
"""
import os

import numpy as np
import pytest
import torch

from oracle import ot_oracle as orc

import _parity

pytestmark = pytest.mark.gpu

RTOL = 1e-4
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT_DIR, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _rel(P, Pref, elem_rtol=_parity.ELEM_RTOL):
    """max-normalised plan error; the elementwise error on entries >= 1e-6 * max is asserted (and both are
    recorded for gpurun_out/parity_report.json) inside tests/_parity.py"""
    import inspect
    return _parity.rel(P, Pref, what=inspect.stack()[1].function, elem_rtol=elem_rtol)


def _dev(x, dev, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=dtype)


# ---------------------------------------------------------------------------
# cost construction
# ---------------------------------------------------------------------------
def test_cost_sqeuclid_golden(cuda_dev, golden_dir):
    from b200ot import ops
    g = _load(golden_dir, "c1_sample_64.npz")
    C = ops.cost_matrix(_dev(g["X"], cuda_dev), _dev(g["Y"], cuda_dev)).cpu().numpy()
    np.testing.assert_allclose(C, g["C"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("n,m,d", [(1, 7, 3), (130, 257, 65), (300, 129, 512)])
def test_cost_ragged_shapes(cuda_dev, n, m, d):
    from b200ot import ops
    rng = np.random.default_rng(n * 1000 + m)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Y = rng.standard_normal((m, d)).astype(np.float32)
    C = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev)).cpu().numpy()
    ref = orc.sqeuclid_cost(X, Y)
    np.testing.assert_allclose(C, ref, rtol=0, atol=1e-5 * max(1.0, np.abs(ref).max()))
    Cc = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev), kind="cosine").cpu().numpy()
    np.testing.assert_allclose(Cc, orc.cosine_cost(X, Y), rtol=0, atol=2e-6)


@pytest.mark.parametrize("n,m,d", [(128, 256, 64), (64, 64, 512), (300, 700, 96), (1000, 1300, 512), (4096, 4096, 512)])
def test_cost_tensor_core_split_is_fp32_accurate(cuda_dev, n, m, d):
    """tcgen05 3-term bf16 split against the float64 oracle (and the plain bf16 product for contrast)."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=n % 5)
    ref = orc.sqeuclid_cost(X, Y)
    xd, yd = _dev(X, cuda_dev), _dev(Y, cuda_dev)
    C6 = ops.cost_matrix(xd, yd, impl="tc", terms=6).cpu().numpy()
    assert np.abs(C6 - ref).max() < 1.5e-6   # fp32-grade: what a float32 FMA loop achieves
    C3 = ops.cost_matrix(xd, yd, impl="tc", terms=3).cpu().numpy()
    assert np.abs(C3 - ref).max() < 1e-5     # 2^-17 per product
    C1 = ops.cost_matrix(xd, yd, impl="tc", terms=1).cpu().numpy()
    assert 1e-5 < np.abs(C1 - ref).max() < 5e-2  # plain bf16: three orders worse, why the split exists
    Cc = ops.cost_matrix(xd, yd, kind="cosine", impl="tc").cpu().numpy()
    assert np.abs(Cc - orc.cosine_cost(X, Y)).max() < 3e-6
    Cs = ops.cost_matrix(xd, yd, impl="simt").cpu().numpy()
    assert np.abs(Cs - ref).max() < 4e-6
    # two fp16 parts per operand (rows scaled by a power of two), 3 products: same grade at half the tensor work
    Ch = ops.cost_matrix(xd, yd, impl="tc", terms="f16").cpu().numpy()
    assert np.abs(Ch - ref).max() < 1.5e-6
    Ch4 = ops.cost_matrix(xd, yd, impl="tc", terms="f16x4").cpu().numpy()
    assert np.abs(Ch4 - ref).max() < 1.5e-6
    Chc = ops.cost_matrix(xd, yd, kind="cosine", impl="tc", terms="f16").cpu().numpy()
    assert np.abs(Chc - orc.cosine_cost(X, Y)).max() < 3e-6


def test_cost_fp16_split_scales_every_row(cuda_dev):
    """The fp16 split must not depend on the magnitude of the data: rows spanning 1e-18 .. 1e+15, a heavy-tailed row
    (one entry 1e4 times the rest), a zero row and an unaligned leading dimension.  Error bound: 2e-6 |x_i||y_j|
    (representation error 2^-23 of the row maximum per entry, worst case sqrt(d) of it in the dot product)."""
    from b200ot import ops
    rng = np.random.default_rng(5)
    n, m, d = 384, 640, 200
    X = rng.standard_normal((n, d))
    Y = rng.standard_normal((m, d))
    X *= 10.0 ** rng.uniform(-18, 15, size=(n, 1))
    Y *= 10.0 ** rng.uniform(-3, 3, size=(m, 1))
    X[7, 3] *= 1e4
    X[11] = 0.0
    Y[5, :] = 0.0
    X = X.astype(np.float32).astype(np.float64)
    Y = Y.astype(np.float32).astype(np.float64)
    xd = torch.zeros(n, d + 5, device=cuda_dev)[:, :d]
    xd.copy_(_dev(X, cuda_dev))
    ref = (X * X).sum(1)[:, None] + (Y * Y).sum(1)[None, :] - 2.0 * X @ Y.T
    bound = np.linalg.norm(X, axis=1)[:, None] * np.linalg.norm(Y, axis=1)[None, :]
    for terms in ("f16", "f16x4", 6):
        Cm = ops.cost_matrix(xd, _dev(Y, cuda_dev), impl="tc", terms=terms).double().cpu().numpy()
        assert np.isfinite(Cm).all()
        # fp32 rounding of |x|^2 + |y|^2 itself is 2^-24 of the larger norm: allow it on top of the product bound
        tol = 2e-6 * bound + 2.5e-7 * ((X * X).sum(1)[:, None] + (Y * Y).sum(1)[None, :]) + 1e-37
        assert (np.abs(Cm - ref) <= tol).all(), (terms, float((np.abs(Cm - ref) / tol).max()))
    # an output with an odd leading dimension takes the scalar store path of the epilogue
    odd = torch.empty(n, m + 1, device=cuda_dev)[:, :m]
    Co = ops.cost_matrix(xd, _dev(Y, cuda_dev), out=odd, impl="tc", terms="f16")
    assert Co.stride(0) == m + 1 and torch.equal(Co, ops.cost_matrix(xd, _dev(Y, cuda_dev), impl="tc", terms="f16"))
    cref = 1.0 - (X / np.maximum(np.linalg.norm(X, axis=1, keepdims=True), 1e-300)) @ \
        (Y / np.maximum(np.linalg.norm(Y, axis=1, keepdims=True), 1e-300)).T
    Cc = ops.cost_matrix(xd, _dev(Y, cuda_dev), kind="cosine", impl="tc", terms="f16").double().cpu().numpy()
    assert np.abs(Cc - cref).max() < 3e-6


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_fot_cost_matches_reference_construction(cuda_dev, golden_dir, impl):
    """b200ot_fot_cost (fp32 FMA) and b200ot_fot_cost_tc (both contractions on tcgen05, 6-term bf16 split) against
    the reference construction (MRI_PET_OT_nojax.py:121-136) and the reference's own init_matrix_np output."""
    from b200ot import ops
    g = _load(golden_dir, "c1_sample_64.npz")
    Ts = np.eye(64) / 64
    M = ops.fot_cost(_dev(g["X"], cuda_dev), _dev(g["Y"], cuda_dev), _dev(Ts, cuda_dev),
                     _dev(Ts.sum(1), cuda_dev), _dev(Ts.sum(0), cuda_dev), impl=impl).cpu().numpy()
    ref = orc.fot_cost_pot(g["X"], g["Y"], Ts)
    np.testing.assert_allclose(M, ref, rtol=0, atol=1e-7 if impl == "simt" else 4e-7 * float(np.abs(ref).max()))
    h = _load(golden_dir, "helpers.npz")  # init_matrix_np of the reference, rectangular
    T = np.random.default_rng(3).random((6, 5))
    Mg = ops.fot_cost(_dev(h["X1"].T, cuda_dev), _dev(h["X2"].T, cuda_dev), _dev(T, cuda_dev),
                      _dev(h["v1"], cuda_dev), _dev(h["v2"], cuda_dev), impl=impl).cpu().numpy()
    np.testing.assert_allclose(Mg, h["constC"] - h["hC1"] @ T @ h["hC2"].T, rtol=0, atol=2e-5)


def test_fot_cost_tensor_core_chain_ragged_shapes(cuda_dev):
    """The tcgen05 chain at the reference-native feature width (d' = 2048) with ragged sample counts, an unaligned
    leading dimension and a dense, non-symmetric sample coupling; float64 NumPy is the reference."""
    from b200ot import ops
    rng = np.random.default_rng(11)
    n, n2, d, d2 = 100, 90, 300, 2048
    A = rng.standard_normal((n, d))
    B = rng.standard_normal((n2, d2)) * 0.7 + 0.2
    T = rng.random((n, n2))
    T /= T.sum()
    w1, w2 = T.sum(1), T.sum(0)
    ref = (A * A).T @ w1[:, None] + ((B * B).T @ w2)[None, :] - 2.0 * A.T @ T @ B
    Ad = torch.zeros(n, d + 3, device=cuda_dev)[:, :d]
    Ad.copy_(_dev(A, cuda_dev))
    scale = float(np.abs(ref).max())
    for impl in ("tc", "simt", "auto"):
        M = ops.fot_cost(Ad, _dev(B, cuda_dev), _dev(T, cuda_dev), _dev(w1, cuda_dev), _dev(w2, cuda_dev),
                         impl=impl).double().cpu().numpy()
        assert M.shape == (d, d2)
        assert float(np.abs(M - ref).max()) <= 2e-6 * scale, impl


# ---------------------------------------------------------------------------
# Sinkhorn: golden vectors from the reference's own sinkhorn_scaling
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("path", ["fused", "robust"])
def test_c1_200_iterations_match_reference(cuda_dev, golden_dir, path):
    import b200ot
    g = _load(golden_dir, "c1_sample_64.npz")
    a = np.ones(64) / 64
    P, lg = b200ot.sinkhorn(a, a, g["C"], float(g["eps"]), numItermax=200, stopThr=0.0, log=True,
                            err_norm="l2sq", path=path, warn=False)
    assert lg["n_iter"] == 200
    assert _rel(P, g["P200"]) < RTOL
    np.testing.assert_allclose(lg["log_u"], np.log(g["u200"]), rtol=0, atol=2e-4)
    np.testing.assert_allclose(lg["log_v"], np.log(g["v200"]), rtol=0, atol=2e-4)
    assert len(lg["err"]) == len(g["err200"])
    # the first check is far above the fp32 noise floor of the squared-L2 error (~1e-13); later ones are not
    np.testing.assert_allclose(lg["err"][:1], g["err200"][:1], rtol=2e-3)
    assert max(lg["err"][1:]) < 1e-12


def test_c1_mirror_rule_same_iteration_count(cuda_dev, golden_dir):
    import b200ot
    g = _load(golden_dir, "c1_sample_64.npz")
    a = np.ones(64) / 64
    K = np.exp(-g["C"] / float(g["eps"]))
    P, lg = b200ot.sinkhorn_scaling(a, a, K, numItermax=2000, stopThr=1e-9, log=True)
    assert lg["n_iter"] == 11 == 10 * (len(g["errconv"]) - 1) + 1
    assert len(lg["err"]) == len(g["errconv"])
    assert _rel(P, g["Pconv"]) < RTOL


def test_c3_cohort_4096_matches_reference(cuda_dev, golden_dir):
    from b200ot import ops
    g = _load(golden_dir, "c3_cohort_4096_summary.npz")
    X, Y = orc.synthetic_embeddings(4096, 4096, 512, config_index=2)
    xd, yd = _dev(X, cuda_dev), _dev(Y, cuda_dev)
    C = ops.cost_matrix(xd, yd)
    a = torch.full((4096,), 1.0 / 4096, device=cuda_dev)
    eps = float(g["eps"])
    f0 = torch.full((4096,), eps * np.log(1.0 / 4096), device=cuda_dev)
    f, gg, info = ops.sinkhorn_potentials(C, a, a, eps, max_iter=200, tol=0.0, err_norm="l2sq", f0=f0)
    assert info["n_iter"] == 200
    np.testing.assert_allclose(f.cpu().numpy() / eps, np.log(g["u200"]), rtol=0, atol=3e-4)
    np.testing.assert_allclose(gg.cpu().numpy() / eps, np.log(g["v200"]), rtol=0, atol=3e-4)
    P = ops.plan(C, f, gg, eps).cpu().numpy()
    assert _rel(P[::512], g["rows200"]) < RTOL
    cost = float(ops.ot_cost(C, f, gg, eps).item())
    assert abs(cost - float(g["cost200"])) < RTOL * abs(float(g["cost200"]))
    bary = ops.apply_plan(C, f, gg, eps, yd, normalise=True).cpu().numpy()
    np.testing.assert_allclose(bary[::512], g["bary_rows"], rtol=0, atol=RTOL * np.abs(g["bary_rows"]).max())
    errs = info["errs"].cpu().numpy()
    assert len(errs) == len(g["err200"])
    np.testing.assert_allclose(errs[:2], g["err200"][:2], rtol=5e-3, atol=1e-13)
    # convergence run, mirror rule: same number of checks and iterations as the reference
    f, gg, info = ops.sinkhorn_potentials(C, a, a, eps, max_iter=2000, tol=1e-9, err_norm="l2sq",
                                          stop_inclusive=True, f0=f0)
    assert info["converged"]
    assert info["n_err"] == len(g["errconv"])
    assert info["n_iter"] == 10 * (len(g["errconv"]) - 1) + 1


# ---------------------------------------------------------------------------
# Sinkhorn: oracle on seeded inputs, both paths, ragged shapes, non-uniform marginals
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(64, 64), (200, 2048), (37, 4100), (1000, 8192), (513, 12288)])
@pytest.mark.parametrize("path", ["fused", "robust"])
def test_paths_match_oracle(cuda_dev, n, m, path):
    """Both single-sweep kernels (256-thread two-per-SM form by default) and the robust kernels."""
    from b200ot import ops
    rng = np.random.default_rng(n + m)
    X, Y = orc.synthetic_embeddings(n, m, 32, config_index=n % 7)
    C = orc.sqeuclid_cost(X, Y)
    a = rng.random(n) + 0.5
    a /= a.sum()
    b = rng.random(m) + 0.5
    b /= b.sum()
    eps = 0.1
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=25, tol=0.0, err_norm="l1", check_every=5,
                                check_phase=0, log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    os.environ["B200OT_RESIDENT"] = "0"  # the launch-per-sweep kernels; tests/test_gpu_resident.py covers the other
    try:
        f, g, info = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=25, tol=0.0,
                                             check_every=5, check_phase=0, err_norm="l1", path=path)
    finally:
        del os.environ["B200OT_RESIDENT"]
    assert info["n_iter"] == 25 and info["status"] == 0
    P = ops.plan(Cd, f, g, eps).cpu().numpy()
    assert _rel(P, Pref) < RTOL
    np.testing.assert_allclose(info["errs"].cpu().numpy(), lg["err"], rtol=2e-2, atol=2e-6)


@pytest.mark.parametrize("n,m", [(5, 3), (33, 31), (2, 1001)])
def test_unaligned_shapes_use_generic_kernels(cuda_dev, n, m):
    from b200ot import ops
    rng = np.random.default_rng(11)
    C = rng.random((n, m))
    a = np.ones(n) / n
    b = np.ones(m) / m
    Pref = orc.sinkhorn_log(C, a, b, 0.2, max_iter=30, tol=0.0)
    f, g, info = ops.sinkhorn_potentials(_dev(C, cuda_dev), _dev(a, cuda_dev), _dev(b, cuda_dev), 0.2,
                                         max_iter=30, tol=0.0)
    P = ops.plan(_dev(C, cuda_dev), f, g, 0.2).cpu().numpy()
    assert _rel(P, Pref) < RTOL


def test_ott_rule_iteration_count(cuda_dev, golden_dir):
    """ott flavour (perturbot/match/fot.py:129-134): max-scaled cost, L1 rule, threshold 1e-3."""
    import b200ot
    g = _load(golden_dir, "fot_ott_512.npz")
    M, _ = orc.fot_cost_ott(g["X"], g["Y"], np.eye(64) / 64)
    Pref, lg = orc.sinkhorn_log_ott(M, float(g["eps"]), log=True)
    out = b200ot.linear_solve(b200ot.Geometry(cost_matrix=M, epsilon=float(g["eps"]), scale_cost="max_cost"),
                              max_iterations=2000)
    assert out.n_iters == lg["n_iter"] and out.converged
    assert _rel(out.matrix, Pref) < RTOL
    Tv, log = b200ot.get_coupling_fot(({0: g["X"]}, {0: g["Y"]}), {0: np.eye(64) / 64}, eps=float(g["eps"]))
    assert _rel(Tv, g["Tv"]) < RTOL
    assert abs(log["cost"][-1] - float(g["cost"])) < RTOL * abs(float(g["cost"]))


def test_feature_coupling_pot_golden(cuda_dev, golden_dir):
    import b200ot
    g = _load(golden_dir, "fot_pot_512.npz")
    # the golden plan was produced by the reference function driving the reference's own in-tree
    # sinkhorn_scaling (squared-L2 rule, inclusive stop): ask the engine for the same rule
    mirror = dict(err_norm="l2sq", stop_inclusive=True)
    Tv, lg = b200ot.get_feature_coupling_pot(({0: g["X"]}, {0: g["Y"]}), {0: np.eye(64) / 64},
                                             eps=float(g["eps"]), **mirror)
    assert lg == {} and Tv.shape == (512, 512) and Tv.dtype == np.float64
    assert _rel(Tv, g["Tv"]) < RTOL
    h = _load(golden_dir, "fot_pot_labels.npz")
    Xd = {1: h["X1"], 0: h["X0"]}
    Yd = {1: h["Y1"], 0: h["Y0"]}
    Tv, _ = b200ot.get_feature_coupling_pot((Xd, Yd), {1: h["Ts1"], 0: h["Ts0"]}, eps=float(h["eps"]), **mirror)
    assert _rel(Tv, h["Tv"]) < RTOL


def test_small_eps_falls_back_and_stays_finite(cuda_dev):
    """eps = 1e-3 on a max-scaled cost: exp(-C/eps) underflows in the reference; the engine must
    return a finite plan with exact row marginals (robust replay of chunks that lose a sum)."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(256, 2048, 16, config_index=5)
    C = orc.sqeuclid_cost(X, Y)
    C = C / C.max()
    a = np.ones(256) / 256
    b = np.ones(2048) / 2048
    eps = 1e-3
    Pref = orc.sinkhorn_log(C, a, b, eps, max_iter=40, tol=0.0)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    f, g, info = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=40, tol=0.0)
    P = ops.plan(Cd, f, g, eps).cpu().numpy()
    assert np.isfinite(P).all() and info["n_iter"] == 40
    np.testing.assert_allclose(P.sum(1), a, rtol=1e-3)
    assert _rel(P, Pref, elem_rtol=None) < 2e-2  # exponents ~1e3: fp32 resolves the plan to ~1e-3 here


def test_warm_start_and_stepper(cuda_dev):
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(512, 4096, 64, config_index=1)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(512) / 512
    b = np.ones(4096) / 4096
    Pref, lg = orc.sinkhorn_log(C, a, b, 0.05, max_iter=30, tol=0.0, log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    st = ops.SinkhornStepper(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.05, max_iter=30, tol=0.0)
    st.enqueue(12)
    assert st.flags()["it"] == 12
    st.enqueue(100)  # more than remain: the rule stops it at max_iter
    fl = st.flags()
    assert fl["it"] == 30 and fl["done"] == 1
    f, g, info = st.finish()
    assert _rel(ops.plan(Cd, f, g, 0.05).cpu().numpy(), Pref) < RTOL
    # warm start from the solution: one more iteration changes nothing
    f2, g2, _ = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.05, max_iter=1, tol=0.0,
                                        f0=f, g0=g)
    Pw = orc.sinkhorn_log(C, a, b, 0.05, max_iter=1, tol=0.0, f0=lg["f"], g0=lg["g"])
    assert _rel(ops.plan(Cd, f2, g2, 0.05).cpu().numpy(), Pw) < RTOL


# ---------------------------------------------------------------------------
# batched small problems (float64 kernel-domain, POT arithmetic)
# ---------------------------------------------------------------------------
def test_batched_matches_reference_loop(cuda_dev, golden_dir):
    from b200ot import ops
    g = _load(golden_dir, "c1_sample_64.npz")
    B = 48
    Xs, Ys, Cs = [g["X"]], [g["Y"]], [g["C"]]
    for i in range(1, B):
        X, Y = orc.synthetic_embeddings(64, 64, 512, config_index=100 + i)
        Xs.append(X)
        Ys.append(Y)
        Cs.append(orc.sqeuclid_cost(X, Y))
    a = np.ones(64) / 64
    eps = float(g["eps"])
    ad = _dev(a, cuda_dev)
    # from embeddings, POT rule (L2, strict), convergence run
    P, lg = ops.sinkhorn_batched(ad, ad, eps, X=_dev(np.stack(Xs), cuda_dev), Y=_dev(np.stack(Ys), cuda_dev),
                                 max_iter=2000, tol=1e-9, err_norm="l2")
    P = P.cpu().numpy()
    n_iter = lg["n_iter"].cpu().numpy()
    for i in range(B):
        Pref, rl = orc.sinkhorn_knopp(a, a, M=Cs[i], reg=eps, numItermax=2000, stopThr=1e-9, err_norm="l2",
                                      log=True)
        assert n_iter[i] == rl["n_iter"], i
        assert _rel(P[i], Pref) < 1e-6
    # from a cost tensor, mirror rule, fixed 200 iterations: the reference's own golden plan
    P, lg = ops.sinkhorn_batched(ad, ad, eps, C3=_dev(np.stack(Cs), cuda_dev), max_iter=200, tol=0.0,
                                 err_norm="l2sq", stop_inclusive=False)
    assert int(lg["n_iter"][0]) == 200
    assert _rel(P[0].cpu().numpy(), g["P200"]) < 1e-6
    np.testing.assert_allclose(lg["u"][0].cpu().numpy(), g["u200"], rtol=1e-6)


def test_batched_numerical_guard(cuda_dev):
    """Underflowing kernel: POT restores the previous iterate and stops (utils.py:55-79)."""
    from b200ot import ops
    rng = np.random.default_rng(5)
    C = rng.random((4, 32, 32)) * 1000.0
    a = np.ones(32) / 32
    P, lg = ops.sinkhorn_batched(_dev(a, cuda_dev), _dev(a, cuda_dev), 1e-3, C3=_dev(C, cuda_dev), max_iter=100,
                                 tol=1e-9)
    for i in range(4):
        Pref, rl = orc.sinkhorn_knopp(a, a, M=C[i].astype(np.float32).astype(np.float64), reg=1e-3,
                                      numItermax=100, stopThr=1e-9, log=True)
        assert int(lg["n_iter"][i]) == rl["n_iter"]
        np.testing.assert_allclose(P[i].cpu().numpy(), Pref, rtol=1e-5, atol=1e-30)


# ---------------------------------------------------------------------------
# epilogues
# ---------------------------------------------------------------------------
def test_apply_plan_and_losses(cuda_dev):
    from b200ot import ops
    rng = np.random.default_rng(9)
    n, m, dv = 150, 260, 70
    C = rng.random((n, m))
    f = rng.standard_normal(n) * 0.05
    g = rng.standard_normal(m) * 0.05
    eps = 0.3
    P = orc.plan_from_potentials(C, f, g, eps)
    V = rng.standard_normal((m, dv))
    U = rng.standard_normal((n, dv))
    Cd, fd, gd = _dev(C, cuda_dev), _dev(f, cuda_dev), _dev(g, cuda_dev)
    np.testing.assert_allclose(ops.apply_plan(Cd, fd, gd, eps, _dev(V, cuda_dev)).cpu().numpy(), P @ V,
                               rtol=0, atol=RTOL * np.abs(P @ V).max())
    np.testing.assert_allclose(ops.apply_plan(Cd, fd, gd, eps, _dev(V, cuda_dev), normalise=True).cpu().numpy(),
                               orc.barycentric(P, V), rtol=0, atol=RTOL * np.abs(V).max())
    np.testing.assert_allclose(ops.apply_plan(Cd, fd, gd, eps, _dev(U, cuda_dev), transpose=True).cpu().numpy(),
                               P.T @ U, rtol=0, atol=RTOL * np.abs(P.T @ U).max())
    assert abs(float(ops.ot_cost(Cd, fd, gd, eps).item()) - orc.ot_cost(P, C)) < RTOL * orc.ot_cost(P, C)
    A = rng.standard_normal((33, 512))
    B = rng.standard_normal((33, 512))
    assert abs(float(ops.cosine_loss(_dev(A, cuda_dev), _dev(B, cuda_dev)).item()) - orc.cosine_loss(A, B)) < 1e-5


# ---------------------------------------------------------------------------
# BASELINE full size: size-independent properties
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [16384, 65536])
def test_full_size_marginal_properties(cuda_dev, n):
    """At sizes the oracle cannot reach: after an f update the row marginals are exact, the reported
    error equals the column-marginal violation of the returned plan, and the error decreases."""
    from b200ot import ops
    m, d = n, 512
    gen = torch.Generator(device="cpu").manual_seed(20251118 + 3)
    X = torch.randn(n, d, generator=gen)
    Y = torch.randn(m, d, generator=gen) + 0.5 * torch.randn(1, d, generator=gen)
    X = (X / X.norm(dim=1, keepdim=True)).to(cuda_dev)
    Y = (Y / Y.norm(dim=1, keepdim=True)).to(cuda_dev)
    C = ops.cost_matrix(X, Y)
    a = torch.full((n,), 1.0 / n, device=cuda_dev)
    eps = 0.05
    f, g, info = ops.sinkhorn_potentials(C, a, a, eps, max_iter=21, tol=0.0, check_every=10, check_phase=1,
                                         err_norm="l1")
    assert info["n_iter"] == 21 and info["status"] == 0 and info["n_err"] == 3
    errs = info["errs"].cpu().numpy()
    assert errs[1] < errs[0] and errs[2] < max(1.05 * errs[1], 5e-6)  # later checks sit on the fp32 floor
    ones = torch.ones((m, 1), device=cuda_dev)
    rows = ops.apply_plan(C, f, g, eps, ones).reshape(-1)
    assert float((rows * n - 1).abs().max()) < 5e-4
    cols = ops.apply_plan(C, f, g, eps, torch.ones((n, 1), device=cuda_dev), transpose=True).reshape(-1)
    l1 = float((cols.double() - 1.0 / m).abs().sum())
    # the checker's own column sums (n fp32 adds per column) carry ~1e-6 relative noise, i.e. ~1e-6 in L1
    assert abs(l1 - errs[2]) < 0.02 * errs[2] + 5e-6
    # fused single-sweep and two-sweep robust paths agree
    f2, g2, _ = ops.sinkhorn_potentials(C, a, a, eps, max_iter=3, tol=0.0, path="robust")
    f3, g3, _ = ops.sinkhorn_potentials(C, a, a, eps, max_iter=3, tol=0.0, path="fused")
    assert float((f2 - f3).abs().max()) / eps < 2e-4 and float((g2 - g3).abs().max()) / eps < 2e-4


def test_pipelined_variant_in_subprocess(cuda_dev):
    """The 512-thread software-pipelined single-sweep kernel (fallback for rows wider than 65536 columns)
    is selected per process with B200OT_FUSED_VARIANT=pipe: run it in a child and compare with the oracle."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path[:0] = [%r, %r]
from oracle import ot_oracle as orc
from b200ot import ops
assert "pipelined" in ops.describe_kernel(300, 12288)
X, Y = orc.synthetic_embeddings(300, 12288, 32, config_index=3)
C = orc.sqeuclid_cost(X, Y); a = np.ones(300) / 300; b = np.ones(12288) / 12288
Pref = orc.sinkhorn_log(C, a, b, 0.1, max_iter=20, tol=0.0)
dev = torch.device("cuda", 0)
Cd = ops.aligned_copy(torch.tensor(C, dtype=torch.float32, device=dev))
f, g, info = ops.sinkhorn_potentials(Cd, torch.tensor(a, dtype=torch.float32, device=dev),
                                     torch.tensor(b, dtype=torch.float32, device=dev), 0.1, max_iter=20, tol=0.0, path="fused")
P = ops.plan(Cd, f, g, 0.1).cpu().numpy()
rel = np.abs(P - Pref).max() / Pref.max()
assert info["n_iter"] == 20 and rel < 1e-4, rel
print("ok", rel)
''' % (ROOT_DIR, PKG_DIR)
    env = dict(os.environ, B200OT_FUSED_VARIANT="pipe")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("n,m,shards", [(301, 2048, 2), (1000, 8192, 4)])
def test_row_sharded_kernels_emulated_on_one_gpu(cuda_dev, n, m, shards):
    """The row-sharded C-ABI entry points (setup / shard_prologue / shard_sweep / shard_finalize), with the
    NCCL all-reduce replaced by a sum over shard objects living on one GPU: same plan as the unsharded oracle,
    every shard takes the same stopping decision."""
    from b200ot import ops, sharded
    X, Y = orc.synthetic_embeddings(n, m, 24, config_index=9)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    eps = 0.1
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=60, tol=1e-4, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    bd = _dev(b, cuda_dev)
    prm = ops.make_params(eps, 60, 1e-4, 10, 0, "l1", False, "auto")
    ks = []
    for r in range(shards):
        lo, hi = sharded.row_range(n, shards, r)
        ks.append(sharded.CudaShardKernels(Cd[lo:hi], _dev(a[lo:hi], cuda_dev), bd, prm))
    for k in ks:
        k.setup()
    tot = sum(k.prologue().clone() for k in ks)
    for k in ks:
        k.finalize(tot, True)
    for _ in range(60):
        tot = sum(k.sweep().clone() for k in ks)
        for k in ks:
            k.finalize(tot, False)
    outs = [k.finish() for k in ks]
    assert len({o[2]["n_iter"] for o in outs}) == 1 and outs[0][2]["n_iter"] == lg["n_iter"]
    assert all(o[2]["converged"] == lg["converged"] for o in outs)
    f = torch.cat([o[0] for o in outs])
    for o in outs[1:]:
        assert torch.equal(o[1], outs[0][1])  # g is replicated bit for bit
    P = ops.plan(Cd, f, outs[0][1], eps).cpu().numpy()
    assert _rel(P, Pref) < RTOL


@pytest.mark.parametrize("n,m,shards", [(301, 2048, 2), (1000, 8192, 4), (520, 1000, 8)])
def test_peer_exchange_kernels_emulated_on_one_gpu(cuda_dev, n, m, shards):
    """The NCCL-free sharded loop (shard_push -> tagged words in every peer's buffer -> shard_finalize_peer polls
    its own buffer), with the peers' buffers living on one GPU: all pushes of an exchange are queued before the
    finalizes that poll for them.  Same plan and iteration count as the oracle, g replicated bit for bit, equal to
    the all-reduce form bit for bit (same fold order), and a second solve on the same buffers (new epoch) works."""
    from b200ot import ops, sharded
    X, Y = orc.synthetic_embeddings(n, m, 24, config_index=9)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    eps = 0.1
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=60, tol=1e-4, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    bd = _dev(b, cuda_dev)
    prm = ops.make_params(eps, 60, 1e-4, 10, 0, "l1", False, "auto")
    nbytes = sharded.PeerExchange.nbytes(shards, m)
    assert nbytes == 2 * shards * ((m + 63) // 64 * 64) * 8
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=cuda_dev) for _ in range(shards)]
    ks, peers = [], []
    for r in range(shards):
        lo, hi = sharded.row_range(n, shards, r)
        ks.append(sharded.CudaShardKernels(Cd[lo:hi], _dev(a[lo:hi], cuda_dev), bd, prm))
        peers.append(sharded.PeerExchange(m, local_bufs=bufs, rank=r))
    g_first = None
    for solve in range(2):
        for k, pe in zip(ks, peers):
            k.setup()
            pe.next_epoch()
        for k, pe in zip(ks, peers):
            k.push(pe, True)
        for k, pe in zip(ks, peers):
            k.finalize_peer(pe, True)
        for _ in range(60):
            for k, pe in zip(ks, peers):
                k.push(pe, False)
            for k, pe in zip(ks, peers):
                k.finalize_peer(pe, False)
        outs = [k.finish() for k in ks]
        assert len({o[2]["n_iter"] for o in outs}) == 1 and outs[0][2]["n_iter"] == lg["n_iter"]
        assert all(o[2]["converged"] == lg["converged"] and o[2]["status"] == 0 for o in outs)
        for o in outs[1:]:
            assert torch.equal(o[1], outs[0][1])  # g is replicated bit for bit
        f = torch.cat([o[0] for o in outs])
        P = ops.plan(Cd, f, outs[0][1], eps).cpu().numpy()
        assert _rel(P, Pref) < RTOL
        if g_first is None:
            g_first = outs[0][1].clone()
        else:
            assert torch.equal(g_first, outs[0][1])
    # the all-reduce form of the same loop, summing the shards in rank order
    for k in ks:
        k.setup()
    tot = ks[0].prologue().clone()
    for k in ks[1:]:
        tot = tot + k.prologue()
    for k in ks:
        k.finalize(tot, True)
    for _ in range(60):
        tot = ks[0].sweep().clone()
        for k in ks[1:]:
            tot = tot + k.sweep()
        for k in ks:
            k.finalize(tot, False)
    g_allreduce = ks[0].finish()[1]
    # the peer form sends the first g update as log-domain words (safe for any eps), the all-reduce form adds linear
    # sums: same value, different rounding
    torch.testing.assert_close(g_allreduce, g_first, rtol=0, atol=1e-5)


def test_cuda_graph_replay_equals_eager_launches(cuda_dev):
    """SinkhornStepper.build_graph / run: replaying a captured 10-iteration graph (and past the stopping
    rule) gives bit-identical potentials to eager launches."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(700, 4096, 48, config_index=8)
    Cd = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev), impl="simt")
    a = torch.full((700,), 1.0 / 700, device=cuda_dev)
    b = torch.full((4096,), 1.0 / 4096, device=cuda_dev)
    st = ops.SinkhornStepper(Cd, a, b, 0.05, max_iter=37, tol=0.0)
    st.enqueue(37)
    f0, g0, i0 = st.finish()
    st2 = ops.SinkhornStepper(Cd, a, b, 0.05, max_iter=37, tol=0.0)
    st2.build_graph(10)
    st2.run(60)  # 6 replays: the last 23 iterations' kernels see done = 1 and do nothing
    f1, g1, i1 = st2.finish()
    assert i0["n_iter"] == i1["n_iter"] == 37
    assert torch.equal(f0, f1) and torch.equal(g0, g1)


def test_c_driven_shard_loop_single_shard(cuda_dev):
    """b200ot_sinkhorn_shard_start / shard_run with comm = NULL (the loop the multi-GPU driver queues from C,
    minus the collective) equals the ordinary solve."""
    from b200ot import ops, sharded
    X, Y = orc.synthetic_embeddings(500, 4096, 24, config_index=11)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(500) / 500
    b = np.ones(4096) / 4096
    Pref, lg = orc.sinkhorn_log(C, a, b, 0.1, max_iter=40, tol=1e-5, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    prm = ops.make_params(0.1, 40, 1e-5, 10, 0, "l1", False, "auto")
    k = sharded.CudaShardKernels(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), prm)
    k.setup()
    k.start_c(None)
    k.run_c(25, None)
    k.run_c(25, None)  # past max_iter / convergence: no-ops
    f, g, info = k.finish()
    assert info["n_iter"] == lg["n_iter"] and info["converged"] == lg["converged"]
    assert _rel(ops.plan(Cd, f, g, 0.1).cpu().numpy(), Pref) < RTOL


@pytest.mark.parametrize("n,m,d,panel_bytes", [(1000, 2048, 64, 1 << 20), (300, 4096, 512, 2 << 20)])
def test_online_solver_matches_streaming_and_oracle(cuda_dev, n, m, d, panel_bytes):
    """C-free solver: cost panels rebuilt on tcgen05 every iteration, never materialised as a whole."""
    from b200ot import ops
    from b200ot.online import OnlineSinkhorn, choose_path
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=12)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    eps = 0.05
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=60, tol=1e-6, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    xd, yd = _dev(X, cuda_dev), _dev(Y, cuda_dev)
    sol = OnlineSinkhorn(xd, yd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=60, tol=1e-6, check_every=10,
                         check_phase=0, err_norm="l1", panel_bytes=panel_bytes)
    assert sol.panel_rows < n  # several panels per iteration
    f, g, info = sol.solve()
    assert info["n_iter"] == lg["n_iter"] and info["converged"] == lg["converged"] and info["status"] == 0
    Cd = ops.cost_matrix(xd, yd, impl="tc")
    P = ops.plan(Cd, f, g, eps).cpu().numpy()
    assert _rel(P, Pref) < RTOL
    fs, gs, _ = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=60, tol=1e-6,
                                        check_every=10, check_phase=0, err_norm="l1")
    assert float((fs - f).abs().max()) / eps < 1e-4 and float((gs - g).abs().max()) / eps < 1e-4
    assert choose_path(65536, 65536, 512, 180 << 30) == "streaming"
    assert choose_path(300000, 300000, 512, 180 << 30) == "online"


def test_fp32_floor_stop_instead_of_spinning_to_numItermax(cuda_dev, golden_dir):
    """POT's default stopThr = 1e-9 (un-squared L2) is below what fp32 can resolve at 64 x 64: with the floor
    rule the drop-in stops a few checks after the float64 reference converged, flags it (status 1) and returns
    the converged plan; with floor_patience = 0 it keeps the pure reference rule and runs to numItermax."""
    import b200ot
    g = _load(golden_dir, "c1_sample_64.npz")
    a = np.ones(64) / 64
    eps = float(g["eps"])
    Pref, rl = orc.sinkhorn_knopp(a, a, M=g["C"], reg=eps, numItermax=2000, stopThr=1e-9, err_norm="l2", log=True)
    P, lg = b200ot.sinkhorn(a, a, g["C"], eps, numItermax=2000, stopThr=1e-9, log=True, warn=False)
    assert lg["status"] == 1 and lg["converged"]
    assert rl["n_iter"] <= lg["n_iter"] <= rl["n_iter"] + 60
    assert _rel(P, Pref) < RTOL
    P0, lg0 = b200ot.sinkhorn(a, a, g["C"], eps, numItermax=300, stopThr=1e-9, log=True, warn=False, floor_patience=0)
    assert lg0["n_iter"] == 300 and lg0["status"] == 0 and not lg0["converged"]
    # a threshold above the floor: identical iteration count to the float64 reference, floor rule never fires
    Pref2, rl2 = orc.sinkhorn_knopp(a, a, M=g["C"], reg=eps, numItermax=2000, stopThr=1e-6, err_norm="l2", log=True)
    P2, lg2 = b200ot.sinkhorn(a, a, g["C"], eps, numItermax=2000, stopThr=1e-6, log=True, warn=False)
    assert lg2["n_iter"] == rl2["n_iter"] and lg2["status"] == 0


def test_foscttm_and_label_grouping(cuda_dev):
    """Plan-quality metric on device (perturbot/perturbot/eval/utils.py:18-45) and the label bucketing of the
    per-epoch coupling (MRI_PET_OT_OT_per_epoch_attn.py:918-937) on device tensors."""
    import b200ot
    rng = np.random.default_rng(21)
    n, d = 300, 24
    true = rng.standard_normal((n, d))
    pred = true + 0.8 * rng.standard_normal((n, d))
    pred[7] = pred[3]          # exact tie handling
    true[7] = true[3]
    ref = np.array(orc.foscttm(pred, true))
    got = np.array(b200ot.foscttm(pred.astype(np.float32), true.astype(np.float32)))
    assert np.abs(got - ref).max() <= 1.0 / (n - 1) + 1e-9  # at most one near-tie flips between fp32 and fp64
    assert abs(got.mean() - ref.mean()) < 2e-4
    P = np.abs(rng.standard_normal((n, n)))
    P[5] = 0.0
    vals, agg = b200ot.get_FOSCTTM(P, pred, true)
    refv = orc.foscttm(orc.barycentric(P, true), true)
    assert abs(agg - float(np.mean(refv))) < 2e-3
    labels = rng.integers(0, 3, size=n)
    feats = rng.standard_normal((n, 8)).astype(np.float32)
    ref_g = orc.group_features_by_label(labels, feats, max_samples_per_label=64)
    dev_g = b200ot.group_features_by_label(torch.tensor(labels, device=cuda_dev), torch.tensor(feats, device=cuda_dev), 64)
    assert list(dev_g.keys()) == list(ref_g.keys())
    for k in ref_g:
        assert dev_g[k].is_cuda and np.array_equal(dev_g[k].cpu().numpy(), ref_g[k])


def test_entropic_gromov_wasserstein_per_label(cuda_dev):
    """get_coupling_egw_ott_fixed (MRI_PET_OT_OT_per_epoch_attn.py:129-186): all labels in one launch, one CTA per
    label, against the float64 restatement of ott's GromovWasserstein: same outer / inner iteration counts, same
    couplings; ragged label sizes, different widths on the two sides, NumPy in -> NumPy out."""
    import b200ot
    rng = np.random.default_rng(5)
    X = np.zeros((64, 40), dtype=np.float32)  # a 6-dimensional cloud ...
    X[:, :6] = rng.standard_normal((64, 6))
    Q, _ = np.linalg.qr(rng.standard_normal((6, 6)))
    perm = rng.permutation(64)
    Y = np.zeros((64, 48), dtype=np.float32)  # ... and an isometric copy of it: rotated, permuted, slightly noisy
    Y[:, :6] = (X[:, :6] @ Q)[perm] + 0.01 * rng.standard_normal((64, 6))
    Xd = {0: X, 2: rng.standard_normal((33, 40)).astype(np.float32), 1: rng.standard_normal((7, 40)).astype(np.float32)}
    Yd = {0: Y, 2: (2 * rng.standard_normal((50, 48))).astype(np.float32),
          1: rng.standard_normal((9, 48)).astype(np.float32)}
    Ts, log = b200ot.get_coupling_egw_ott_fixed((Xd, Yd), eps=5e-3)
    Tref, lref = orc.get_coupling_egw_ott_fixed((Xd, Yd), eps=5e-3)
    assert list(Ts.keys()) == [0, 2, 1]
    for l in Xd:
        assert isinstance(Ts[l], np.ndarray) and Ts[l].shape == Tref[l].shape
        assert log[l]["n_iters_outer"] == lref[l]["n_iters_outer"], (l, log[l], lref[l])
        assert log[l]["inner_iterations"] == lref[l]["inner_iterations"]
        assert log[l]["converged_outer"] == lref[l]["converged_outer"]
        assert log[l]["converged_inner"] == lref[l]["converged_inner"]
        assert abs(log[l]["GW cost"] - lref[l]["GW cost"]) < 1e-6 * max(1.0, abs(lref[l]["GW cost"]))
        assert _rel(Ts[l], Tref[l]) < RTOL
    assert (Tref[0].argmax(1) == np.argsort(perm)).mean() > 0.9  # the isometric copy is identified ...
    assert (Ts[0].argmax(1) == Tref[0].argmax(1)).all()           # ... by both
    # CUDA tensors in -> couplings stay on the device
    Tt, _ = b200ot.get_coupling_egw_ott_fixed(({0: _dev(Xd[1], cuda_dev)}, {0: _dev(Yd[1], cuda_dev)}), eps=5e-2)
    assert Tt[0].is_cuda and Tt[0].shape == (7, 9)
    np.testing.assert_allclose(Tt[0].double().cpu().numpy(), orc.egw_ott(Xd[1], Yd[1], eps=5e-2)[0], rtol=0,
                               atol=RTOL * float(Tt[0].max()))


def test_label_constrained_coot_bcd_matches_reference(cuda_dev, golden_dir):
    """cotl_numpy(algo="sinkhorn", algo2="sinkhorn") (perturbot/perturbot/match/cot_labels.py:14-225) against the
    golden produced by the reference's own function.  The reference's exit rule (|cost_old - cost| < 1e-7 on a cost
    of ~17) is below fp32 resolution, so the number of BCD rounds may differ by a few; costs, the feature coupling
    and the sample couplings must agree."""
    import b200ot
    g = _load(golden_dir, "cotl_sinkhorn.npz")
    keys = [int(k) for k in g["keys"]]
    Xd = {k: g[f"X{k}"] for k in keys}
    Yd = {k: g[f"Y{k}"] for k in keys}
    Ts, Tv, cost, lg = b200ot.cotl_numpy(Xd, Yd, niter=2000, algo="sinkhorn", reg=float(g["reg"]), algo2="sinkhorn",
                                         reg2=float(g["reg"]), verbose=False, log=True)
    assert list(Ts.keys()) == keys and Tv.shape == g["Tv"].shape and Tv.dtype == np.float64
    assert 3 <= len(lg["cost"]) <= 4 * len(g["costs"])
    assert abs(cost - float(g["cost"])) < 1e-3 * abs(float(g["cost"]))
    k0 = min(len(lg["cost"]), len(g["costs"]), 10)
    np.testing.assert_allclose(lg["cost"][:k0], g["costs"][:k0], rtol=1e-4)  # the early rounds are far from the floor
    assert _rel(Tv, g["Tv"], elem_rtol=None) < 5e-3
    for k in keys:
        assert _rel(Ts[k], g[f"Ts{k}"], elem_rtol=None) < 5e-3
    Ts2, log2 = b200ot.get_coupling_cotl_sinkhorn((Xd, Yd), eps=float(g["reg"]))
    assert set(Ts2) == set(keys) and "time" in log2 and len(log2["cost"]) == len(lg["cost"])
    with pytest.raises(b200ot.B200OTError):
        b200ot.cotl_numpy(Xd, Yd)  # the default 'emd' variants are POT's network simplex


def test_per_epoch_coupling_pipeline_stays_on_device(cuda_dev):
    """compute_pet_to_mri_coupling (MRI_PET_OT_OT_per_epoch_attn.py:940-960) after feature_extract: label buckets
    -> per-label entropic GW (PET -> MRI) -> feature coupling on the block-diagonal, with CUDA tensors end to end,
    against the same composition of the oracle's restatements."""
    import b200ot
    rng = np.random.default_rng(8)
    N, d = 150, 32
    labels = rng.integers(0, 3, size=N)
    mri = np.abs(rng.standard_normal((N, d))).astype(np.float32)
    pet = np.abs(rng.standard_normal((N, d)) + 0.3 * labels[:, None]).astype(np.float32)
    T = b200ot.compute_pet_to_mri_coupling(_dev(mri, cuda_dev), _dev(pet, cuda_dev),
                                           torch.as_tensor(labels, device=cuda_dev), max_samples_per_label=40)
    assert T.is_cuda and T.shape == (d, d)
    gm = orc.group_features_by_label(labels, mri, max_samples_per_label=40)
    gp = orc.group_features_by_label(labels, pet, max_samples_per_label=40)
    Td, _ = orc.get_coupling_egw_ott_fixed((gp, gm))
    Tref, _ = orc.get_coupling_fot((gp, gm), Td)
    assert _rel(T.double().cpu().numpy(), Tref) < RTOL
    # NumPy in -> NumPy out, like the reference
    Tn = b200ot.compute_pet_to_mri_coupling(mri, pet, labels, max_samples_per_label=40)
    assert isinstance(Tn, np.ndarray) and _rel(Tn, Tref) < RTOL


def test_eot_and_egw_default_entry_points(cuda_dev):
    """get_coupling_eot_ott / get_coupling_egw_ott (perturbot/perturbot/match/ott_egwl.py:129-206,299-372)."""
    import b200ot
    rng = np.random.default_rng(12)
    Xd = {1: rng.standard_normal((30, 20)).astype(np.float32), 0: rng.standard_normal((45, 20)).astype(np.float32)}
    Yd = {1: (rng.standard_normal((28, 20)) + 0.5).astype(np.float32), 0: rng.standard_normal((40, 20)).astype(np.float32)}
    T, log = b200ot.get_coupling_eot_ott((Xd, Yd), eps=5e-2)
    X = np.concatenate([Xd[1], Xd[0]])
    Y = np.concatenate([Yd[1], Yd[0]])
    C = orc.sqeuclid_cost(X, Y)
    Pref, lg = orc.sinkhorn_log_ott(C, 5e-2, log=True)
    assert T.shape == (75, 68) and log["n_iters_outer"] == lg["n_iter"] and log["converged"] == lg["converged"]
    assert _rel(T, Pref) < RTOL
    dual = float(lg["f"].mean() + lg["g"].mean())
    assert abs(log["OT cost"] - dual) < 1e-4 * max(1.0, abs(dual))
    Ts, lgw = b200ot.get_coupling_egw_ott(({0: Xd[1]}, {0: Yd[1]}), eps=5e-2)
    Tr, lr = orc.egw_ott(Xd[1], Yd[1], eps=5e-2, gw_max_iterations=1000)
    assert lgw[0]["n_iters_outer"] == lr["n_iters_outer"] and _rel(Ts[0], Tr) < RTOL


@pytest.mark.parametrize("n,m", [(520, 4096), (301, 12288)])
def test_fused_iteration_with_peer_words_single_rank(cuda_dev, n, m):
    """shard_run_peer at world = 1 with B200OT_FUSE=1: ONE persistent launch for the queued iterations (sweep + fold
    + tagged-word push + poll + finalize + state machine in the sweep's own cooperative cluster launch).  With a single rank the push and the poll hit the same
    buffer, so the whole exchange protocol runs on one GPU without any kernel waiting for another launch.
    Same plan and iteration count as the oracle; equal to the separate-launch form bit for bit."""
    from b200ot import ops, sharded
    from b200ot import _lib as _lib_mod
    X, Y = orc.synthetic_embeddings(n, m, 24, config_index=10)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    eps = 0.1
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=60, tol=1e-4, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    prm = ops.make_params(eps, 60, 1e-4, 10, 0, "l1", False, "auto")
    buf = torch.zeros(sharded.PeerExchange.nbytes(1, m), dtype=torch.uint8, device=cuda_dev)
    outs = []
    lib = _lib_mod.load()
    launches0 = lib.b200ot_sinkhorn_counter(0)
    for fuse in ("1", "0", "tail"):
        os.environ["B200OT_RESIDENT"] = "0"
        # the persistent fused form is opt-in; "tail" = the default of run_peer: plain sweep + ONE fold / push / poll /
        # finalize launch per iteration; "0" = the three separate launches
        os.environ["B200OT_FUSE"] = "1" if fuse == "1" else "0"
        try:
            k = sharded.CudaShardKernels(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), prm)
            pe = sharded.PeerExchange(m, local_bufs=[buf], rank=0)
            pe.epoch = {"1": 7, "0": 8, "tail": 9}[fuse]
            k.setup()
            k.push(pe, True)
            k.finalize_peer(pe, True)
            if fuse in ("1", "tail"):
                k.run_peer(35, pe)   # fused launches / sweep + merged tail
                k.run_peer(35, pe)   # past convergence: no-ops
            else:
                for _ in range(60):  # the separate-launch form of the same loop
                    k.push(pe, False)
                    k.finalize_peer(pe, False)
            outs.append(k.finish())
        finally:
            del os.environ["B200OT_RESIDENT"]
            del os.environ["B200OT_FUSE"]
    f, g, info = outs[0]
    assert info["n_iter"] == lg["n_iter"] and info["converged"] == lg["converged"] and info["status"] == 0
    assert _rel(ops.plan(Cd, f, g, eps).cpu().numpy(), Pref) < RTOL
    np.testing.assert_allclose(info["errs"].cpu().numpy(), lg["err"], rtol=2e-2, atol=2e-6)
    for other in outs[1:]:  # same fold order in all three forms
        assert torch.equal(other[0], f) and torch.equal(other[1], g)
        assert other[2]["n_iter"] == info["n_iter"]
    # the fused form really ran (a refused cooperative cluster launch would silently fall back)
    assert lib.b200ot_sinkhorn_counter(1) == 0, lib.b200ot_last_cuda_error()
    assert lib.b200ot_sinkhorn_counter(0) - launches0 >= 1


def test_measured_sweep_rate_feeds_the_balanced_split(cuda_dev):
    """sharded.measure_sweep_rate times the local sweep of a set-up shard (rows per ms, state not advancing);
    balanced_bounds turns the ranks' rates into row ranges.  One GPU: two shards of the same problem measured in
    turn must give comparable rates, an (almost) even split, and the solver state must be untouched."""
    from b200ot import ops, sharded
    n, m = 1024, 8192
    X, Y = orc.synthetic_embeddings(n, m, 32, config_index=6)
    Cd = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev))
    b = torch.full((m,), 1.0 / m, device=cuda_dev)
    prm = ops.make_params(0.05, 100, 0.0, 10, 1, "l2", False, "auto")
    rates = []
    os.environ["B200OT_RESIDENT"] = "0"
    try:
        for lo, hi in (sharded.row_range(n, 2, 0), sharded.row_range(n, 2, 1)):
            k = sharded.CudaShardKernels(Cd[lo:hi], torch.full((hi - lo,), 1.0 / n, device=cuda_dev), b, prm)
            k.setup()
            k.finalize(k.prologue(), True)
            rates.append(sharded.measure_sweep_rate(k, sweeps=50, warm=10))
            assert k.flags()["it"] == 0 and k.flags()["bad"] == 0
    finally:
        del os.environ["B200OT_RESIDENT"]
    assert all(r > 0 for r in rates) and max(rates) / min(rates) < 1.5
    bounds = sharded.balanced_bounds(n, rates)
    assert bounds[0][0] == 0 and bounds[1][1] == n and bounds[0][1] == bounds[1][0]


def test_sharded_solve_recovers_from_a_lost_sum(cuda_dev):
    """eps = 1e-3 on a max-scaled cost: the single-sweep kernel loses row sums in the first iterations.  The
    sharded driver must do what b200ot_sinkhorn_solve does: rewind to the chunk snapshot and replay on the robust
    kernels (the sharded path had no recovery)."""
    from b200ot import ops, sharded
    X, Y = orc.synthetic_embeddings(256, 2048, 16, config_index=5)
    C = orc.sqeuclid_cost(X, Y)
    C = C / C.max()
    a = np.ones(256) / 256
    b = np.ones(2048) / 2048
    eps = 1e-3
    Pref = orc.sinkhorn_log(C, a, b, eps, max_iter=40, tol=0.0)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    buf = torch.zeros(sharded.PeerExchange.nbytes(1, 2048), dtype=torch.uint8, device=cuda_dev)
    os.environ["B200OT_RESIDENT"] = "0"
    try:
        # the peer-exchange loop (default for N > 1): first g update and robust iterations travel as log-domain words
        for peer in (sharded.PeerExchange(2048, local_bufs=[buf], rank=0),):
            f, g, info = sharded.solve_sharded(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=40, tol=0.0,
                                               peer=peer)
            assert info["n_iter"] == 40 and info["status"] == 0
            P = ops.plan(Cd, f, g, eps).cpu().numpy()
            assert np.isfinite(P).all()
            np.testing.assert_allclose(P.sum(1), a, rtol=1e-3)
            assert _rel(P, Pref, elem_rtol=None) < 2e-2
    finally:
        del os.environ["B200OT_RESIDENT"]
