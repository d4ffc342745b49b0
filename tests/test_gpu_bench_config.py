"""Parity of the BENCHMARKED kernel configurations against the float64 oracle.

bench.py times n = m = 65536 (BASELINE configs[3]) on ``sweep_lite_kernel<8>``: clusters of 8 CTAs, 32 columns
per thread, rows streamed through L2 with the evict-first policy.  The kernel configuration depends on the row
WIDTH (m) and on whether C fits L2, not on the number of rows, so a 2048 x 65536 problem runs exactly the
benchmarked instantiation (asserted on ``describe_kernel``) at a size the oracle finishes in seconds.  The
inputs are the first 2048 rows of the bench workload itself (bench.synthetic_rows, seed 20251118 + 3).

The oracle here is the kernel-domain float64 Sinkhorn-Knopp that is pinned bit for bit against the reference's
``sinkhorn_scaling`` (perturbot/perturbot/match/utils.py:6-115, tests/test_oracle.py); with C in [0, 4] and
eps = 0.05, K >= e^-80 has not underflowed, so the kernel- and log-domain iterations are the same map
(SURVEY appendix A).  Tolerances: plan 1e-4 (max-normalised) and the elementwise bound of tests/_parity.py,
potentials 2e-4 in units of eps, iteration counts exactly equal under the mirror and the ott rule.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ot_oracle as orc

import _parity

pytestmark = pytest.mark.gpu

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = _parity.RTOL
EPS = 0.05


def _bench_rows(n_rows, n_total, m):
    sys.path.insert(0, ROOT_DIR)
    import bench
    X, Y = bench.synthetic_rows(n_total, m, 0, n_rows, 20251118 + 3, None)
    return X.numpy(), Y.numpy()


def _dev(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=torch.float32)


def _check_solve(cuda_dev, n, m, expect_kernel, fixed_iters, d=512):
    from b200ot import ops
    if d == 512:
        X, Y = _bench_rows(n, 65536 if m == 65536 else m, m)
    else:  # low-dimensional clouds: C spreads over [0, 4] and Sinkhorn needs tens of iterations
        X, Y = orc.synthetic_embeddings(n, m, d, config_index=3)
    desc = ops.describe_kernel(n, m)
    for token in expect_kernel:
        assert token in desc, (token, desc)
    C = orc.sqeuclid_cost(X, Y)
    a = np.full(n, 1.0 / n)
    b = np.full(m, 1.0 / m)
    Cd = ops.cost_matrix(_dev(X, cuda_dev), _dev(Y, cuda_dev))          # the bench's own cost path (tcgen05)
    assert float(np.abs(Cd[::64].double().cpu().numpy() - C[::64]).max()) < 2e-6
    ad, bd = _dev(a, cuda_dev), _dev(b, cuda_dev)
    f0 = torch.full((n,), EPS * float(np.log(1.0 / n)), device=cuda_dev)    # POT / mirror start u = 1/n

    # ---- fixed iteration count (what bench.py times), mirror error rule recorded every 10 iterations
    Pref, lg = orc.sinkhorn_knopp(a, b, M=C, reg=EPS, numItermax=fixed_iters, stopThr=0.0, err_norm="l2sq", log=True)
    f, g, info = ops.sinkhorn_potentials(Cd, ad, bd, EPS, max_iter=fixed_iters, tol=0.0, err_norm="l2sq", path="fused",
                                         f0=f0)
    assert info["n_iter"] == fixed_iters == lg["n_iter"] and info["status"] == 0
    assert info["n_err"] == len(lg["err"])
    np.testing.assert_allclose(f.double().cpu().numpy() / EPS, np.log(lg["u"]), rtol=0, atol=2e-4)
    np.testing.assert_allclose(g.double().cpu().numpy() / EPS, np.log(lg["v"]), rtol=0, atol=2e-4)
    P = ops.plan(Cd, f, g, EPS).cpu().numpy()
    assert _parity.rel(P, Pref, what=f"bench_config_{n}x{m}_fixed{fixed_iters}") < RTOL
    del P, Pref
    # the first check is far above the fp32 floor of the squared-L2 error
    np.testing.assert_allclose(info["errs"].cpu().numpy()[:1], lg["err"][:1], rtol=5e-3)

    # ---- convergence, mirror rule (squared L2 <= 1e-9 every 10 iterations, perturbot/match/utils.py:48,80-89)
    _, lm = orc.sinkhorn_knopp(a, b, M=C, reg=EPS, numItermax=2000, stopThr=1e-9, err_norm="l2sq", log=True)
    _, _, im = ops.sinkhorn_potentials(Cd, ad, bd, EPS, max_iter=2000, tol=1e-9, err_norm="l2sq", stop_inclusive=True,
                                       path="fused", f0=f0)
    assert im["converged"] and im["n_iter"] == lm["n_iter"], (im["n_iter"], lm["n_iter"])
    assert im["n_err"] == len(lm["err"])

    # ---- convergence, ott rule (L1 < 1e-3 every 10 iterations, zero start potentials, fot.py:129-134)
    Po, lo = orc.sinkhorn_knopp(a, b, M=C, reg=EPS, numItermax=2000, stopThr=1e-3, err_norm="l1", check_phase=0,
                                u0=np.ones(n), v0=np.ones(m), log=True)
    fo, go, io = ops.sinkhorn_potentials(Cd, ad, bd, EPS, max_iter=2000, tol=1e-3, err_norm="l1", check_phase=0,
                                         path="fused")
    assert io["converged"] and io["n_iter"] == lo["n_iter"], (io["n_iter"], lo["n_iter"])
    P = ops.plan(Cd, fo, go, EPS).cpu().numpy()
    assert _parity.rel(P, Po, what=f"bench_config_{n}x{m}_ott_rule") < RTOL
    return desc


@pytest.mark.parametrize("d", [512, 8])
def test_benchmarked_sweep_lite_q8_configuration_matches_oracle(cuda_dev, d):
    """n = 2048 rows of the bench workload (d = 512; converges in 11 / 10 iterations like every high-dimensional
    cloud), and a d = 8 cloud on the same kernel that needs 31 (mirror rule) / 40 (ott rule) iterations;
    m = 65536: the exact kernel string bench.py reports."""
    from b200ot import ops
    os.environ.pop("B200OT_RESIDENT", None)
    desc = _check_solve(cuda_dev, 2048, 65536,
                        ["sweep_lite_kernel", "cluster=8 CTAs x 256 threads", "32 cols/thread", "evict-first"], 50, d=d)
    # identical to the 65536-row configuration bench.py runs (the row count only bounds the number of clusters)
    assert desc == ops.describe_kernel(65536, 65536)
    # two CTAs per SM: 33 clusters of 8 on a B200 (a kernel that grows past 113 KB of shared memory per CTA, static
    # included, silently drops to one CTA per SM and half the bandwidth)
    import re
    nclusters = int(re.search(r"(\d+) clusters", desc).group(1))
    assert nclusters * 8 > torch.cuda.get_device_properties(cuda_dev).multi_processor_count, desc


def test_resident_512_thread_variant_at_8192_columns_matches_oracle(cuda_dev):
    """rows wider than 4096 columns take the one-CTA-per-SM, 512-thread form of the resident kernel."""
    os.environ.pop("B200OT_RESIDENT", None)
    _check_solve(cuda_dev, 4096, 8192, ["resident_kernel", "x 512 threads"], 60)
