"""GPU parity of the resident (persistent, cooperative) Sinkhorn kernel, csrc/resident.cu.

The kernel serves the short-iteration problems (n*m <= 8192^2): BASELINE configs C1 / C3 and the
reference-native feature problems of MRI_PET_OT_nojax.py:91-145.  It is checked against the float64 oracle,
against the launch-per-sweep kernels on the same inputs (same iteration counts, same error history), for
both ring modes (shared-memory resident / streamed with alternating sweep direction) and for the cases the
state machine has to get right inside the kernel: early stop, max_iter inside a launch, continuation across
launches, the numerical fallback and CUDA-graph capture.
"""
import os
from contextlib import contextmanager

import numpy as np
import pytest
import torch

from oracle import ot_oracle as orc

import _parity

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@contextmanager
def _env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    try:
        for k, v in kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _dev(x, dev, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=dev, dtype=dtype)


def _rel(P, Pref, elem_rtol=_parity.ELEM_RTOL):
    """max-normalised plan error; the elementwise error on entries >= 1e-6 * max is asserted (and both are
    recorded for gpurun_out/parity_report.json) inside tests/_parity.py"""
    import inspect
    return _parity.rel(P, Pref, what=inspect.stack()[1].function, elem_rtol=elem_rtol)


def _problem(n, m, seed, d=32):
    rng = np.random.default_rng(seed)
    X, Y = orc.synthetic_embeddings(n, m, d, config_index=seed % 7)
    C = orc.sqeuclid_cost(X, Y)
    a = rng.random(n) + 0.5
    a /= a.sum()
    b = rng.random(m) + 0.5
    b /= b.sum()
    return C, a, b


def test_describe_reports_the_resident_kernel(cuda_dev):
    from b200ot import ops
    with _env(B200OT_RESIDENT=None):
        assert "resident_kernel" in ops.describe_kernel(4096, 4096)
        assert "shared-memory resident" in ops.describe_kernel(2048, 2048)
        assert "snake" in ops.describe_kernel(8192, 8192)
        assert "resident_kernel" not in ops.describe_kernel(16384, 16384)
        assert "resident_kernel" not in ops.describe_kernel(512, 12288)  # rows wider than one CTA covers
    with _env(B200OT_RESIDENT=0):
        assert "resident_kernel" not in ops.describe_kernel(4096, 4096)


# shapes: tiny, fewer groups than SMs, ragged last group, every (threads, quads) instantiation, masked last quad,
# shared-memory resident and streamed blocks
@pytest.mark.parametrize("n,m", [(7, 12), (64, 64), (100, 260), (333, 1024), (512, 512), (1001, 2048), (2048, 2048),
                                 (700, 3000), (3000, 4096), (900, 6148), (1000, 8192), (40000, 64)])
def test_resident_matches_oracle_and_per_sweep_kernels(cuda_dev, n, m):
    from b200ot import ops
    C, a, b = _problem(n, m, n + m)
    eps = 0.1
    Pref, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=25, tol=0.0, err_norm="l1", check_every=5, check_phase=0,
                                log=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    ad, bd = _dev(a, cuda_dev), _dev(b, cuda_dev)
    kw = dict(max_iter=25, tol=0.0, check_every=5, check_phase=0, err_norm="l1", path="fused")
    with _env(B200OT_RESIDENT=1):
        assert "resident_kernel" in ops.describe_kernel(n, m)
        f, g, info = ops.sinkhorn_potentials(Cd, ad, bd, eps, **kw)
    assert info["n_iter"] == 25 and info["status"] == 0 and info["n_err"] == 5
    P = ops.plan(Cd, f, g, eps).cpu().numpy()
    assert _rel(P, Pref) < RTOL
    np.testing.assert_allclose(info["errs"].cpu().numpy(), lg["err"], rtol=2e-2, atol=2e-6)
    with _env(B200OT_RESIDENT=0):
        f0, g0, info0 = ops.sinkhorn_potentials(Cd, ad, bd, eps, **kw)
    # same arithmetic per element, different summation order: potentials agree to fp32 rounding (units of eps)
    np.testing.assert_allclose(f.cpu().numpy(), f0.cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(g.cpu().numpy(), g0.cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(info["errs"].cpu().numpy(), info0["errs"].cpu().numpy(), rtol=1e-2, atol=1e-6)


@pytest.mark.parametrize("n,m", [(3000, 4096), (1000, 8192)])
def test_snake_and_forward_sweeps_agree_and_runs_are_bit_reproducible(cuda_dev, n, m):
    from b200ot import ops
    C, a, b = _problem(n, m, 5)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    ad, bd = _dev(a, cuda_dev), _dev(b, cuda_dev)
    out = {}
    for snake in (1, 0):
        with _env(B200OT_RESIDENT=1, B200OT_RES_SNAKE=snake):
            assert ("snake" in ops.describe_kernel(n, m)) == bool(snake)
            out[snake] = [ops.sinkhorn_potentials(Cd, ad, bd, 0.05, max_iter=31, tol=0.0, path="fused")
                          for _ in range(2)]
    for snake in (1, 0):
        (f1, g1, i1), (f2, g2, i2) = out[snake]
        assert torch.equal(f1, f2) and torch.equal(g1, g2)  # fixed-order folds: no atomics on data
        assert torch.equal(i1["errs"], i2["errs"])
    np.testing.assert_allclose(out[1][0][0].cpu().numpy(), out[0][0][0].cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(out[1][0][1].cpu().numpy(), out[0][0][1].cpu().numpy(), rtol=0, atol=1e-4)


@pytest.mark.parametrize("n,m,rule", [(512, 512, "ott"), (2048, 2048, "mirror"), (4096, 4096, "ott"),
                                      (4096, 4096, "mirror")])
def test_stopping_rule_fires_on_the_same_iteration(cuda_dev, n, m, rule):
    """Iteration counts: resident == launch-per-sweep == float64 oracle (BASELINE north star)."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(n, m, 64, config_index=3)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    eps = 0.05
    if rule == "ott":
        kw = dict(tol=1e-3, err_norm="l1", check_every=10, check_phase=0)
    else:
        kw = dict(tol=1e-9, err_norm="l2sq", check_every=10, check_phase=1, stop_inclusive=True)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    ad, bd = _dev(a, cuda_dev), _dev(b, cuda_dev)
    res = {}
    for flag in (1, 0):
        with _env(B200OT_RESIDENT=flag):
            res[flag] = ops.sinkhorn_potentials(Cd, ad, bd, eps, max_iter=2000, path="fused", **kw)
    i1, i0 = res[1][2], res[0][2]
    assert i1["converged"] and i0["converged"]
    assert i1["n_iter"] == i0["n_iter"] and i1["n_err"] == i0["n_err"]
    np.testing.assert_allclose(i1["errs"].cpu().numpy(), i0["errs"].cpu().numpy(), rtol=2e-2, atol=1e-12)
    if n <= 2048:
        _, lg = orc.sinkhorn_log(C, a, b, eps, max_iter=2000, log=True, **kw)
        assert i1["n_iter"] == lg["n_iter"]
        P = ops.plan(Cd, res[1][0], res[1][1], eps).cpu().numpy()
        Pref = np.exp((lg["f"][:, None] + lg["g"][None, :] - C) / eps)
        assert _rel(P, Pref) < RTOL


def test_continuation_across_launches_and_budget_past_max_iter(cuda_dev):
    from b200ot import ops
    C, a, b = _problem(600, 4096, 9)
    Pref = orc.sinkhorn_log(C, a, b, 0.05, max_iter=30, tol=0.0)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    with _env(B200OT_RESIDENT=1):
        st = ops.SinkhornStepper(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.05, max_iter=30, tol=0.0)
        st.enqueue(1)
        st.enqueue(11)
        assert st.flags()["it"] == 12
        st.enqueue(100)  # more than remain: the kernel stops itself at max_iter
        fl = st.flags()
        assert fl["it"] == 30 and fl["done"] == 1
        st.enqueue(5)  # done: the launch returns at once
        assert st.flags()["it"] == 30
        f, g, info = st.finish()
    assert info["n_err"] == 3
    assert _rel(ops.plan(Cd, f, g, 0.05).cpu().numpy(), Pref) < RTOL


def test_resident_mixes_with_per_sweep_launches_on_one_workspace(cuda_dev):
    """The two kernel families share the state block: a solve may switch between them at any iteration."""
    from b200ot import ops
    C, a, b = _problem(1500, 2048, 21)
    Pref = orc.sinkhorn_log(C, a, b, 0.05, max_iter=24, tol=0.0)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    st = ops.SinkhornStepper(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.05, max_iter=24, tol=0.0)
    for i, flag in enumerate((1, 0, 1, 0)):
        with _env(B200OT_RESIDENT=flag):
            st.enqueue(6)
        assert st.flags()["it"] == 6 * (i + 1)
    f, g, _ = st.finish()
    assert _rel(ops.plan(Cd, f, g, 0.05).cpu().numpy(), Pref) < RTOL


def test_lost_sum_inside_the_resident_kernel_replays_robustly(cuda_dev):
    """eps small enough that a column sum vanishes in fp32: the kernel raises `bad` on every CTA at the same
    iteration, the host rewinds to the snapshot and finishes on the running-max kernels."""
    from b200ot import ops
    X, Y = orc.synthetic_embeddings(256, 2048, 16, config_index=5)
    C = orc.sqeuclid_cost(X, Y)
    C = C / C.max()
    a = np.ones(256) / 256
    b = np.ones(2048) / 2048
    eps = 1e-3
    Pref = orc.sinkhorn_log(C, a, b, eps, max_iter=40, tol=0.0)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    with _env(B200OT_RESIDENT=1):
        f, g, info = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), eps, max_iter=40, tol=0.0)
    P = ops.plan(Cd, f, g, eps).cpu().numpy()
    assert np.isfinite(P).all() and info["n_iter"] == 40
    np.testing.assert_allclose(P.sum(1), a, rtol=1e-3)
    assert _rel(P, Pref) < 2e-2


def test_zero_mass_rows_and_columns(cuda_dev):
    """a_i = 0 / b_j = 0 entries carry no mass (potential -inf), as in the launch-per-sweep kernels."""
    from b200ot import ops
    C, a, b = _problem(300, 1024, 4)
    a[::7] = 0.0
    a /= a.sum()
    b[5::11] = 0.0
    b /= b.sum()
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    res = {}
    for flag in (1, 0):
        with _env(B200OT_RESIDENT=flag):
            f, g, info = ops.sinkhorn_potentials(Cd, _dev(a, cuda_dev), _dev(b, cuda_dev), 0.1, max_iter=20,
                                                 tol=0.0, path="fused")
            res[flag] = (ops.plan(Cd, f, g, 0.1).cpu().numpy(), info)
    P1, P0 = res[1][0], res[0][0]
    assert res[1][1]["n_iter"] == res[0][1]["n_iter"] == 20
    assert np.isfinite(P1).all()
    assert np.all(P1[::7] == 0) and np.all(P1[:, 5::11] == 0)
    np.testing.assert_allclose(P1, P0, rtol=0, atol=RTOL * P0.max())
    np.testing.assert_allclose(P1.sum(1), a, rtol=1e-4, atol=1e-9)


def test_cuda_graph_of_a_resident_launch(cuda_dev):
    from b200ot import ops
    C, a, b = _problem(2048, 2048, 2)
    Cd = ops.aligned_copy(_dev(C, cuda_dev))
    ad, bd = _dev(a, cuda_dev), _dev(b, cuda_dev)
    with _env(B200OT_RESIDENT=1):
        st = ops.SinkhornStepper(Cd, ad, bd, 0.05, max_iter=40, tol=0.0)
        st.enqueue(40)
        f_e, g_e, _ = st.finish()
        st2 = ops.SinkhornStepper(Cd, ad, bd, 0.05, max_iter=40, tol=0.0)
        st2.build_graph(10)
        st2.run(40)
        assert st2.flags()["it"] == 40
        f_g, g_g, _ = st2.finish()
    # four launches of 10 iterations vs one of 40: identical arithmetic in the same order
    assert torch.equal(f_e, f_g) and torch.equal(g_e, g_g)
