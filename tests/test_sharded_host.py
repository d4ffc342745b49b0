"""Host-side logic of the row-sharded solver (b200ot/sharded.py) on CPU with gloo, world_size 2.

The CUDA kernels cannot run here, so the per-rank kernel interface is filled by a float64 NumPy
emulation of what the C-ABI kernels compute (test infrastructure); what is under test is the
orchestration: the row partition, the one all-reduce per iteration, the replicated stopping rule
(every rank must stop at the same iteration), and that the sharded result equals the unsharded oracle.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ot_oracle as orc


def test_balanced_bounds_follow_the_measured_rates():
    """Rows proportional to the ranks' sweep rates, contiguous, in groups of 4, covering everything; equal rates
    reproduce the even split."""
    from b200ot.sharded import balanced_bounds, row_range
    n = 65536
    assert balanced_bounds(n, [1.0] * 8) == [row_range(n, 8, r) for r in range(8)]
    rates = [341.0, 350.0, 365.0, 345.0, 350.0, 352.0, 348.0, 360.0]
    b = balanced_bounds(n, rates)
    assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(7))
    assert all((hi - lo) % 4 == 0 for lo, hi in b)
    times = [(hi - lo) / r for (lo, hi), r in zip(b, rates)]
    assert max(times) / min(times) < 1.002                       # every rank finishes its sweep together
    even = [8192 / r for r in rates]
    assert max(times) < max(even) * 0.975                         # 7 % spread: the slowest rank gated the even split (-2.9 %)
    for n2, w in ((100, [1, 2, 3]), (4096, [1, 1]), (12, [5, 1, 1])):
        bb = balanced_bounds(n2, w)
        assert bb[0][0] == 0 and bb[-1][1] == n2 and all(bb[i][1] == bb[i + 1][0] for i in range(len(w) - 1))
        assert all(hi > lo for lo, hi in bb)


def test_row_range_partitions_everything():
    from b200ot.sharded import row_range
    for n in (1, 3, 4, 7, 64, 65, 1000, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [row_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in spans:
                assert hi == lo or lo % 4 == 0  # non-empty shards start on a group-of-4 row boundary
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 7 or n < 4 * world  # one group of 4 plus a ragged tail


class NumpyShardKernels:
    """float64 emulation of the per-rank kernels (setup / prologue / sweep / finalize / flags / finish)."""

    def __init__(self, C_local, a_local, b, eps, max_iter, tol, check_every, check_phase):
        self.C, self.a, self.b, self.eps = C_local, a_local, b, eps
        self.max_iter, self.tol, self.ce, self.cp = max_iter, tol, check_every, check_phase

    def setup(self):
        self.f = np.zeros(len(self.a))
        self.g = np.zeros(len(self.b))
        self.it, self.done, self.converged, self.errs = 0, 0, 0, []

    def prologue(self):
        return torch.from_numpy(np.exp((self.f[:, None] + self.g[None, :] - self.C) / self.eps).sum(0))

    def sweep(self):
        if self.done:
            return torch.zeros(len(self.b), dtype=torch.float64)
        t = np.exp((self.f[:, None] + self.g[None, :] - self.C) / self.eps)
        w = self.a / t.sum(1)
        self.f = self.f + self.eps * np.log(w)
        return torch.from_numpy((t * w[:, None]).sum(0))

    def finalize(self, s_total, is_prologue):
        if self.done:
            return
        s = s_total.numpy()
        g_next = self.g + self.eps * (np.log(self.b) - np.log(s))
        if is_prologue:
            self.g = g_next
            return
        self.it += 1
        if self.it % self.ce == self.cp % self.ce:
            err = float(np.abs(s - self.b).sum())
            self.errs.append(err)
            if err < self.tol:
                self.converged = self.done = 1
                return
        if self.it >= self.max_iter:
            self.done = 1
            return
        self.g = g_next

    def flags(self):
        return {"it": self.it, "done": self.done, "converged": self.converged, "bad": 0, "n_err": len(self.errs)}

    def finish(self):
        return self.f, self.g, {"n_iter": self.it, "converged": bool(self.converged), "errs": self.errs}


def _worker(rank, world, port, n, m, tol, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from b200ot.sharded import ShardedSinkhorn, row_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    lo, hi = row_range(n, world, rank)
    k = NumpyShardKernels(C[lo:hi], a[lo:hi], b, 0.1, 200, tol, 10, 0)
    drv = ShardedSinkhorn(k)
    f, g, info = drv.solve(200, check_every=10, check_phase=0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), f=f, g=g, lo=lo, hi=hi, n_iter=info["n_iter"],
             converged=info["converged"], allreduces=drv.allreduces, errs=np.array(info["errs"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("tol", [1e-3, 0.0])
def test_two_rank_gloo_solve_matches_unsharded_oracle(tmp_path, tol):
    n, m, world = 50, 36, 2
    port = 29500 + (os.getpid() % 500) + (0 if tol else 1)
    mp.spawn(_worker, args=(world, port, n, m, tol, str(tmp_path)), nprocs=world, join=True)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    Pref, lg = orc.sinkhorn_log(C, a, b, 0.1, max_iter=200, tol=tol, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # every rank took the same decisions
    assert len({int(p["n_iter"]) for p in parts}) == 1
    assert int(parts[0]["n_iter"]) == lg["n_iter"]
    assert bool(parts[0]["converged"]) == lg["converged"]
    # one all-reduce for the first g update + one per iteration actually queued
    assert int(parts[0]["allreduces"]) >= 1 + lg["n_iter"]
    f = np.concatenate([p["f"] for p in parts])
    np.testing.assert_allclose(parts[0]["g"], parts[1]["g"], rtol=0, atol=0)  # replicated, bit-identical
    P = orc.plan_from_potentials(C, f, parts[0]["g"], 0.1)
    np.testing.assert_allclose(P, Pref, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(parts[0]["errs"], lg["err"], rtol=1e-6, atol=1e-13)  # float64 noise floor


# ---------------------------------------------------------------------------
# peer-exchange form of the loop (no collective call per iteration)
# ---------------------------------------------------------------------------
class FilePeerExchange:
    """Stand-in for sharded.PeerExchange on CPU: the ranks' exchange buffers are one shared memory-mapped file;
    words are (value, tag) pairs, slabs [buffer owner][parity][source rank][column], like the CUDA layout."""

    def __init__(self, path, world, rank, m):
        self.world, self.rank, self.m = world, rank, m
        self.val = np.memmap(path + ".val", dtype=np.float64, mode="r+", shape=(world, 2, world, m))
        self.tag = np.memmap(path + ".tag", dtype=np.int64, mode="r+", shape=(world, 2, world, m))
        self.epoch = 0

    def next_epoch(self):
        self.epoch = (self.epoch + 1) & 0xFFF
        return self.epoch


class NumpyPeerKernels(NumpyShardKernels):
    """The push / finalize_peer / run_peer entry points of CudaShardKernels, restated over FilePeerExchange."""

    def _x(self, is_prologue):
        return 0 if is_prologue else self.it + 1

    def push(self, peer, is_prologue):
        if self.done:
            return
        s = (self.prologue() if is_prologue else self.sweep()).numpy()
        x = self._x(is_prologue)
        tag = (peer.epoch << 20) | (x + 1)
        for dst in range(peer.world):
            peer.val[dst, x & 1, peer.rank, :] = s
            peer.tag[dst, x & 1, peer.rank, :] = tag  # value first, tag second: the tag publishes the value

    def finalize_peer(self, peer, is_prologue):
        if self.done:
            return
        x = self._x(is_prologue)
        tag = (peer.epoch << 20) | (x + 1)
        import time as _t
        t0 = _t.time()
        while not (peer.tag[peer.rank, x & 1] == tag).all():
            assert _t.time() - t0 < 60, "a peer never pushed"
            _t.sleep(0.0005)
        tot = np.zeros(peer.m)
        for r in range(peer.world):  # rank order: identical bits on every rank
            tot = tot + np.array(peer.val[peer.rank, x & 1, r, :])
        self.finalize(torch.from_numpy(tot), is_prologue)

    def run_peer(self, iters, peer):
        for _ in range(iters):
            self.push(peer, False)
            self.finalize_peer(peer, False)


def _peer_worker(rank, world, port, n, m, tol, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from b200ot.sharded import ShardedSinkhorn, row_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    lo, hi = row_range(n, world, rank)
    peer = FilePeerExchange(os.path.join(out_dir, "xbuf"), world, rank, m)
    results = []
    for solve in range(2):  # the second solve reuses the buffers under a new epoch
        k = NumpyPeerKernels(C[lo:hi], a[lo:hi], b, 0.1, 200, tol, 10, 0)
        drv = ShardedSinkhorn(k, peer=peer)
        assert drv.peer is peer and drv.c_loop
        f, g, info = drv.solve(200, check_every=10, check_phase=0)
        assert drv.allreduces == 0  # no collective in the loop
        results.append((f, g, info))
        dist.barrier()
    assert peer.epoch == 2
    (f, g, info), (f2, g2, info2) = results
    assert np.array_equal(g, g2) and np.array_equal(f, f2) and info["n_iter"] == info2["n_iter"]
    np.savez(os.path.join(out_dir, f"peer{rank}.npz"), f=f, g=g, n_iter=info["n_iter"], converged=info["converged"])
    dist.destroy_process_group()


@pytest.mark.parametrize("tol", [1e-3, 0.0])
def test_two_rank_peer_exchange_loop_matches_unsharded_oracle(tmp_path, tol):
    """ShardedSinkhorn with a PeerExchange: start() bumps the epoch and pushes the prologue, run() queues run_peer
    chunks, every rank folds the slabs in rank order and takes the same decisions; no all-reduce is issued."""
    n, m, world = 50, 36, 2
    for suffix, dt in ((".val", np.float64), (".tag", np.int64)):
        np.memmap(str(tmp_path / ("xbuf" + suffix)), dtype=dt, mode="w+", shape=(world, 2, world, m)).flush()
    port = 30100 + (os.getpid() % 500) + (0 if tol else 1)
    mp.spawn(_peer_worker, args=(world, port, n, m, tol, str(tmp_path)), nprocs=world, join=True)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    Pref, lg = orc.sinkhorn_log(C, a, b, 0.1, max_iter=200, tol=tol, err_norm="l1", check_every=10, check_phase=0,
                                log=True)
    parts = [np.load(tmp_path / f"peer{r}.npz") for r in range(world)]
    assert int(parts[0]["n_iter"]) == int(parts[1]["n_iter"]) == lg["n_iter"]
    assert bool(parts[0]["converged"]) == lg["converged"]
    np.testing.assert_array_equal(parts[0]["g"], parts[1]["g"])
    f = np.concatenate([p["f"] for p in parts])
    np.testing.assert_allclose(orc.plan_from_potentials(C, f, parts[0]["g"], 0.1), Pref, rtol=1e-9, atol=1e-15)


# ---------------------------------------------------------------------------
# recovery: a lost sum on ONE rank stops every rank, the chunk is rewound and replayed on the robust path
# ---------------------------------------------------------------------------
class FlakyShardKernels(NumpyShardKernels):
    """The fast path of rank `fail_rank` loses a row sum in iteration `fail_it` (what the CUDA sweep reports as
    `bad`).  Like the kernels, the rank then sends NaN column sums, so every rank's finalize sees NaN, raises `bad`
    and stops in the same iteration; snapshot / rewind / use_robust_path are the C-ABI calls of CudaShardKernels."""

    def __init__(self, *a, rank=0, fail_rank=1, fail_it=3, **kw):
        super().__init__(*a, **kw)
        self.rank, self.fail_rank, self.fail_it = rank, fail_rank, fail_it
        self.robust, self.bad, self.replayed_from = False, 0, None

    def setup(self):
        super().setup()
        self.bad = 0

    def sweep(self):
        s = super().sweep()
        if not self.robust and not self.done and self.rank == self.fail_rank and self.it + 1 == self.fail_it:
            return torch.full_like(s, float("nan"))
        return s

    def finalize(self, s_total, is_prologue):
        if not self.done and bool(torch.isnan(s_total).any()):
            self.bad = self.done = 1
            return
        super().finalize(s_total, is_prologue)

    def snapshot(self):
        self.snap = (self.f.copy(), self.g.copy(), self.it, list(self.errs))

    def rewind(self):
        self.f, self.g, self.it, self.errs = self.snap[0].copy(), self.snap[1].copy(), self.snap[2], list(self.snap[3])
        self.done = self.converged = self.bad = 0
        self.replayed_from = self.it

    def use_robust_path(self):
        self.robust = True

    def flags(self):
        fl = super().flags()
        fl["bad"] = self.bad
        return fl


def _worker_flaky(rank, world, port, n, m, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "ot-based-heterogeneous-multi-modal-fusion-embedding-for-ad-analysis-_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from b200ot.sharded import ShardedSinkhorn, row_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    a = np.ones(n) / n
    b = np.ones(m) / m
    lo, hi = row_range(n, world, rank)
    k = FlakyShardKernels(C[lo:hi], a[lo:hi], b, 0.1, 200, 1e-3, 10, 0, rank=rank, fail_rank=1, fail_it=13)
    f, g, info = ShardedSinkhorn(k).solve(200, check_every=10, check_phase=0)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), f=f, g=g, n_iter=info["n_iter"], converged=info["converged"],
             replayed_from=-1 if k.replayed_from is None else k.replayed_from, robust=k.robust, bad=k.bad)
    dist.destroy_process_group()


def test_lost_sum_on_one_rank_rewinds_every_rank(tmp_path):
    n, m, world = 50, 36, 2
    port = 29500 + (os.getpid() % 500) + 7
    mp.spawn(_worker_flaky, args=(world, port, n, m, str(tmp_path)), nprocs=world, join=True)
    X, Y = orc.synthetic_embeddings(n, m, 16, config_index=4)
    C = orc.sqeuclid_cost(X, Y)
    Pref, lg = orc.sinkhorn_log(C, np.ones(n) / n, np.ones(m) / m, 0.1, max_iter=200, tol=1e-3, err_norm="l1",
                                check_every=10, check_phase=0, log=True)
    assert lg["n_iter"] > 13  # the failure happens before convergence
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for p in parts:  # BOTH ranks saw the failure, rewound to the start of the chunk (iteration 10) and switched path
        assert int(p["replayed_from"]) == 10 and bool(p["robust"]) and int(p["bad"]) == 0
        assert int(p["n_iter"]) == lg["n_iter"] and bool(p["converged"]) == lg["converged"]
    f = np.concatenate([p["f"] for p in parts])
    np.testing.assert_allclose(orc.plan_from_potentials(C, f, parts[0]["g"], 0.1), Pref, rtol=1e-9, atol=1e-15)
