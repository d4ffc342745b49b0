"""GPU parity of the single-pass tcgen05 plan-application kernel (csrc/apply_tc.cu) through the C ABI
(b200ot_apply_plan_tc, b200ot_envelope_bwd) against float64 NumPy restatements of the reference's epilogues:
`pet @ T.t()` (MRI_PET_OT_OT_per_epoch_attn.py:728), the barycentric projection `(T / rowsum) @ Y` with
`rowsum == 0 -> 1e-30` (perturbot/perturbot/eval/match.py:202-206) and the envelope gradient (oracle.envelope_grads).
Tolerance: 1e-4 of the largest reference entry (north star: fused embedding within fp32 relative 1e-4)."""
import numpy as np
import pytest
import torch

from oracle import ot_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _dev(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=torch.float32)


def _close(got, ref, what):
    got = got.double().cpu().numpy()
    scale = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max())
    assert err <= RTOL * scale, (what, err / scale)
    return err / scale


def _random_problem(n, m, dv, seed, eps=0.3):
    rng = np.random.default_rng(seed)
    C = rng.random((n, m))
    f = rng.standard_normal(n) * 0.05
    g = rng.standard_normal(m) * 0.05
    V = rng.standard_normal((m, dv))
    U = rng.standard_normal((n, dv))
    return C, f, g, V, U, eps


# shapes: one tile; ragged everything with an unaligned leading dimension (scalar loads); both TMEM halves with
# the K range split over CTAs; a second half narrower than 256; width not a multiple of 16; dv = 512 at C3 size
@pytest.mark.parametrize("n,m,dv", [(128, 2048, 16), (301, 1003, 70), (1024, 4096, 512), (520, 2048, 300),
                                    (2000, 640, 33), (4096, 4096, 512)])
def test_apply_tc_matches_float64(cuda_dev, n, m, dv):
    from b200ot import ops
    C, f, g, V, U, eps = _random_problem(n, m, dv, n + m + dv)
    P = orc.plan_from_potentials(C, f, g, eps)
    Cd, fd, gd = _dev(C, cuda_dev), _dev(f, cuda_dev), _dev(g, cuda_dev)
    Vd, Ud = _dev(V, cuda_dev), _dev(U, cuda_dev)
    Z, rs = ops.apply_plan(Cd, fd, gd, eps, Vd, impl="tc", return_rowsum=True)
    _close(Z, P @ V, "P V")
    _close(rs, P.sum(1), "row sums")
    _close(ops.apply_plan(Cd, fd, gd, eps, Vd, normalise=True, impl="tc"), orc.barycentric(P, V), "barycentric")
    Zt, cs = ops.apply_plan(Cd, fd, gd, eps, Ud, transpose=True, impl="tc", return_rowsum=True)
    _close(Zt, P.T @ U, "P^T U")
    _close(cs, P.sum(0), "column sums")
    _close(ops.apply_plan(Cd, fd, gd, eps, Ud, transpose=True, normalise=True, impl="tc"),
           orc.barycentric(P.T, U), "column-normalised P^T U")
    # the generic SIMT kernel agrees with the tensor-core one to the same tolerance
    _close(ops.apply_plan(Cd, fd, gd, eps, Vd, impl="simt"), P @ V, "P V (simt)")


def test_apply_tc_on_a_converged_plan_with_empty_rows(cuda_dev):
    """eps = 0.05 on normalised embeddings (plan entries span many decades), zero-mass rows (f = -inf):
    the normalised projection of an empty row is 0 like the reference's `marg == 0 -> 1e-30` guard."""
    from b200ot import ops
    n, m, d = 1536, 2048, 64
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.full(n, 1.0 / (n - 3))
    a[[5, 700, 1535]] = 0.0
    b = np.full(m, 1.0 / m)
    Cd = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev))
    f, g, info = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.05, max_iter=40, tol=0.0)
    assert info["status"] == 0 and bool(torch.isinf(f[5])) and float(f[5]) < 0
    P = orc.plan_from_potentials(C, f.double().cpu().numpy(), g.double().cpu().numpy(), 0.05)
    Z = ops.apply_plan(Cd, f, g, 0.05, _dev(Y, cuda_dev), normalise=True, impl="tc")
    ref = orc.barycentric(P, Y)
    _close(Z, ref, "barycentric projection of a converged plan")
    assert float(Z[5].abs().max()) == 0.0 and float(Z[700].abs().max()) == 0.0


@pytest.mark.parametrize("n,m,d", [(1024, 1536, 64), (700, 900, 512), (4096, 4096, 512)])
def test_envelope_backward_single_launch_matches_oracle(cuda_dev, n, m, d):
    """dX = 2 (diag(P1) X - P Y), dY = 2 (diag(P^T 1) Y - P^T X): both halves from ONE launch."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=7)
    C = orc.sqeuclid_cost(X, Y)
    a = np.full(n, 1.0 / n)
    b = np.full(m, 1.0 / m)
    xd, yd = _dev(X, cuda_dev), _dev(Y, cuda_dev)
    Cd = ops.cost_matrix(xd, yd)
    f, g, _ = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.1, max_iter=30, tol=0.0)
    P = orc.plan_from_potentials(C, f.double().cpu().numpy(), g.double().cpu().numpy(), 0.1)
    dX, dY = orc.envelope_grads(X, Y, P)
    gx, gy = ops.envelope_bwd(Cd, f, g, 0.1, xd, yd, impl="tc")
    _close(gx, dX, "dX")
    _close(gy, dY, "dY")
    sx, sy = ops.envelope_bwd(Cd, f, g, 0.1, xd, yd, impl="simt")
    _close(sx, dX, "dX (simt)")
    _close(sy, dY, "dY (simt)")


def test_apply_tc_is_bit_reproducible_and_handles_wide_right_hand_sides(cuda_dev):
    from b200ot import ops
    C, f, g, V, U, eps = _random_problem(777, 1500, 600, 3)  # dv > 512: two slabs of columns
    Cd, fd, gd, Vd = _dev(C, cuda_dev), _dev(f, cuda_dev), _dev(g, cuda_dev), _dev(V, cuda_dev)
    Z1 = ops.apply_plan(Cd, fd, gd, eps, Vd, impl="tc")
    Z2 = ops.apply_plan(Cd, fd, gd, eps, Vd, impl="tc")
    assert torch.equal(Z1, Z2)  # fixed-order folds, no atomics
    _close(Z1, orc.plan_from_potentials(C, f, g, eps) @ V, "dv = 600")
