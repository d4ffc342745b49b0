"""Row-sharded Sinkhorn across the GPUs of one box.

Rank r owns rows ``row_range(n, world, r)`` of the cost matrix (and of ``a``, ``f``);
``b`` and ``g`` are replicated.  The f update is purely local.  The g update needs the
column sums of the plan over *all* rows, so each iteration all-reduces one fp32 vector
of m column partials (256 KiB at m = 65536) over NCCL/NVLink; every rank then runs the
same finalize kernel (marginal error, stopping rule, next g) on identical data, so the
ranks never disagree about convergence and no other communication is needed.

The orchestration is written against a tiny kernel interface (``setup / prologue / sweep /
finalize / flags / finish``) so the host logic can be exercised with gloo on CPU in the
tests; the product implementation of that interface is ``CudaShardKernels`` (C ABI).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import B200OTError, check


def row_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) owned by `rank`: contiguous blocks, sizes differ by at most one group of 4
    rows so every shard keeps the 16-byte row alignment of the single-sweep kernel."""
    groups = (n + 3) // 4
    base, extra = divmod(groups, world)
    lo_g = rank * base + min(rank, extra)
    hi_g = lo_g + base + (1 if rank < extra else 0)
    return min(lo_g * 4, n), min(hi_g * 4, n)


def balanced_bounds(n: int, weights: Sequence[float]) -> List[Tuple[int, int]]:
    """Contiguous row blocks whose sizes are proportional to `weights` (measured sweep rates of the ranks), in
    groups of 4 rows like ``row_range``.  Every iteration of the sharded loop ends in an exchange of all ranks, so
    the loop runs at the pace of the slowest GPU; on a box whose GPUs differ (7 % between the fastest and the
    slowest of eight was measured, same binary, same shard) an even split wastes the difference."""
    world = len(weights)
    w = [max(float(x), 0.0) for x in weights]
    if world < 1 or not sum(w) > 0.0:
        raise B200OTError("balanced_bounds needs positive weights")
    groups = (n + 3) // 4
    total = sum(w)
    cuts, acc = [0], 0.0
    for r in range(world - 1):
        acc += w[r]
        g = int(round(groups * acc / total))
        if groups >= world:  # every rank keeps at least one group
            g = min(max(g, cuts[-1] + 1), groups - (world - 1 - r))
        cuts.append(min(max(g, cuts[-1]), groups))
    cuts.append(groups)
    return [(min(cuts[r] * 4, n), min(cuts[r + 1] * 4, n)) for r in range(world)]


def measure_sweep_rate(kern, sweeps: int = 200, warm: int = 100) -> float:
    """Rows per millisecond of this rank's local sweep (no exchange; the state does not advance), timed with CUDA
    events after `warm` untimed sweeps so that the clocks have settled under load.  `kern` must be set up
    (setup + first g update)."""
    for _ in range(warm):
        kern.sweep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(sweeps):
        kern.sweep()
    e1.record()
    torch.cuda.synchronize()
    return kern.n * sweeps / max(e0.elapsed_time(e1), 1e-6)


class NcclComm:
    """A communicator owned by libb200ot (ncclCommInitRank through the C ABI): rank 0 creates the unique id,
    torch.distributed broadcasts it, every rank joins.  Used so the per-iteration all-reduce can be queued
    from C on the compute stream instead of from Python on PyTorch's NCCL stream."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        idt = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            check(self.lib.b200ot_nccl_unique_id(buf), "b200ot_nccl_unique_id")
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        idt = idt.to(dev)
        dist.broadcast(idt, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(idt.cpu().tolist())
        self.handle = C.c_void_p()
        check(self.lib.b200ot_nccl_init(raw, self.world, self.rank, C.byref(self.handle)), "b200ot_nccl_init")

    def close(self):
        if self.handle:
            self.lib.b200ot_nccl_destroy(self.handle)
            self.handle = C.c_void_p()


class PeerExchange:
    """Exchange buffers for the NCCL-free sharded loop: one buffer per rank, mapped into every peer with CUDA IPC
    (collective constructor: every rank of `group` must call it).  `local_bufs` builds the same object from
    plain device tensors of ONE process instead -- several shards emulated on one GPU, used by the tests."""

    def __init__(self, m: int, group: Optional[dist.ProcessGroup] = None, local_bufs=None, rank: int = 0):
        self.lib = _lib.load()
        self.m = int(m)
        self._opened, self._own = [], None
        if local_bufs is not None:
            self.world, self.rank = len(local_bufs), int(rank)
            self._keep = local_bufs
            ptrs = [t.data_ptr() for t in local_bufs]
        else:
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            nbytes = self.lib.b200ot_peer_exchange_bytes(self.world, self.m)
            own, handle = C.c_void_p(), (C.c_ubyte * 64)()
            check(self.lib.b200ot_peer_alloc(nbytes, C.byref(own), handle), "b200ot_peer_alloc")
            self._own = own
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            ptrs = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    ptrs.append(own.value)
                    continue
                p = C.c_void_p()
                check(self.lib.b200ot_peer_open(hb, C.byref(p)), "b200ot_peer_open")
                self._opened.append(p)
                ptrs.append(p.value)
        self.ptrs = (C.c_void_p * self.world)(*ptrs)
        self.epoch = 0

    @staticmethod
    def nbytes(world: int, m: int) -> int:
        return _lib.load().b200ot_peer_exchange_bytes(int(world), int(m))

    def next_epoch(self) -> int:
        """A new solve on the same buffers: the same sequence of values on every rank."""
        self.epoch = (self.epoch + 1) & 0x7FF
        return self.epoch

    def close(self):
        for p in self._opened:
            self.lib.b200ot_peer_close(p)
        self._opened = []
        if self._own is not None:
            self.lib.b200ot_peer_free(self._own)
            self._own = None


class CudaShardKernels:
    """The C-ABI implementation of the per-rank kernels (include/b200ot.h, row-sharded form)."""

    def __init__(self, C_local: torch.Tensor, a_local: torch.Tensor, b: torch.Tensor, prm, path="auto",
                 f0=None, g0=None):
        self.lib = _lib.load()
        self.C, self.ldc = ops._matrix(C_local, "C_local")
        self.n, self.m = self.C.shape
        self.a = ops._vector(a_local, "a_local", self.n)
        self.b = ops._vector(b, "b", self.m)
        self.prm = prm
        self.path = _lib.PATHS[path]
        self.f0, self.g0 = f0, g0
        dev = self.C.device
        self.ws = torch.empty(self.lib.b200ot_sinkhorn_workspace_bytes(self.n, self.m) + 256,
                              dtype=torch.uint8, device=dev)
        self.s = torch.zeros(self.m, dtype=torch.float32, device=dev)

    @ops._on_self_device
    def setup(self):
        check(self.lib.b200ot_sinkhorn_setup(self.n, self.m, ops._ptr(self.a), ops._ptr(self.b),
                                             ops._ptr(self.f0), ops._ptr(self.g0), C.byref(self.prm),
                                             ops._ws_ptr(self.ws), self.ws.numel() - 256, ops._stream()),
              "b200ot_sinkhorn_setup")

    @ops._on_self_device
    def prologue(self) -> torch.Tensor:
        check(self.lib.b200ot_sinkhorn_shard_prologue(ops._ptr(self.C), self.ldc, self.n, self.m,
                                                      ops._ws_ptr(self.ws), ops._ptr(self.s), ops._stream()),
              "b200ot_sinkhorn_shard_prologue")
        return self.s

    @ops._on_self_device
    def sweep(self) -> torch.Tensor:
        check(self.lib.b200ot_sinkhorn_shard_sweep(ops._ptr(self.C), self.ldc, self.n, self.m, self.path,
                                                   ops._ws_ptr(self.ws), ops._ptr(self.s), ops._stream()),
              "b200ot_sinkhorn_shard_sweep")
        return self.s

    @ops._on_self_device
    def finalize(self, s_total: torch.Tensor, is_prologue: bool):
        check(self.lib.b200ot_sinkhorn_shard_finalize(self.n, self.m, ops._ws_ptr(self.ws), ops._ptr(s_total),
                                                      int(is_prologue), ops._stream()),
              "b200ot_sinkhorn_shard_finalize")

    # -- C-driven loop: everything, including the all-reduce, is queued by one call on the compute stream
    @ops._on_self_device
    def start_c(self, comm: Optional["NcclComm"]):
        check(self.lib.b200ot_sinkhorn_shard_start(ops._ptr(self.C), self.ldc, self.n, self.m, ops._ws_ptr(self.ws),
                                                   ops._ptr(self.s), comm.handle if comm else None, ops._stream()),
              "b200ot_sinkhorn_shard_start")

    @ops._on_self_device
    def run_c(self, iters: int, comm: Optional["NcclComm"]):
        check(self.lib.b200ot_sinkhorn_shard_run(ops._ptr(self.C), self.ldc, self.n, self.m, int(iters), self.path,
                                                 ops._ws_ptr(self.ws), ops._ptr(self.s),
                                                 comm.handle if comm else None, ops._stream()),
              "b200ot_sinkhorn_shard_run")

    # -- peer-memory loop: no collective call, the column sums travel as tagged words over NVLink
    @ops._on_self_device
    def push(self, peer: "PeerExchange", is_prologue: bool):
        check(self.lib.b200ot_sinkhorn_shard_push(ops._ptr(self.C), self.ldc, self.n, self.m, self.path,
                                                  ops._ws_ptr(self.ws), peer.ptrs, peer.world, peer.rank, peer.epoch,
                                                  int(is_prologue), ops._stream()), "b200ot_sinkhorn_shard_push")

    @ops._on_self_device
    def finalize_peer(self, peer: "PeerExchange", is_prologue: bool):
        check(self.lib.b200ot_sinkhorn_shard_finalize_peer(self.n, self.m, ops._ws_ptr(self.ws),
                                                           peer.ptrs[peer.rank], peer.world, peer.epoch,
                                                           int(is_prologue), ops._stream()),
              "b200ot_sinkhorn_shard_finalize_peer")

    @ops._on_self_device
    def run_peer(self, iters: int, peer: "PeerExchange"):
        check(self.lib.b200ot_sinkhorn_shard_run_peer(ops._ptr(self.C), self.ldc, self.n, self.m, int(iters),
                                                      self.path, ops._ws_ptr(self.ws), peer.ptrs, peer.world,
                                                      peer.rank, peer.epoch, ops._stream()),
              "b200ot_sinkhorn_shard_run_peer")

    @ops._on_self_device
    def snapshot(self):
        check(self.lib.b200ot_sinkhorn_snapshot(self.n, self.m, ops._ws_ptr(self.ws), ops._stream()),
              "b200ot_sinkhorn_snapshot")

    @ops._on_self_device
    def rewind(self):
        check(self.lib.b200ot_sinkhorn_rewind(self.n, self.m, ops._ws_ptr(self.ws), ops._stream()),
              "b200ot_sinkhorn_rewind")

    def use_robust_path(self):
        self.path = _lib.PATH_ROBUST

    @ops._on_self_device
    def flags(self) -> dict:
        out = torch.empty(8, dtype=torch.int32, device=self.C.device)
        check(self.lib.b200ot_sinkhorn_peek(ops._ws_ptr(self.ws), ops._ptr(out), ops._stream()),
              "b200ot_sinkhorn_peek")
        v = out.cpu().tolist()
        return {"it": v[0], "done": v[1], "converged": v[2], "bad": v[4], "n_err": v[5]}

    @ops._on_self_device
    def finish(self):
        f = torch.empty(self.n, dtype=torch.float32, device=self.C.device)
        g = torch.empty(self.m, dtype=torch.float32, device=self.C.device)
        res = torch.zeros(8, dtype=torch.int32, device=self.C.device)
        errs = torch.zeros(512, dtype=torch.float32, device=self.C.device)
        check(self.lib.b200ot_sinkhorn_finish(self.n, self.m, ops._ws_ptr(self.ws), ops._ptr(f), ops._ptr(g),
                                              ops._ptr(res), ops._ptr(errs), 512, ops._stream()),
              "b200ot_sinkhorn_finish")
        r = res.cpu()
        n_err = int(r[3])
        return f, g, {"n_iter": int(r[0]), "converged": bool(r[1]), "status": int(r[2]), "n_err": n_err,
                      "err": float(r[4:5].view(torch.float32)[0]), "errs": errs[:min(n_err, 512)]}


class ShardedSinkhorn:
    """Drives one row shard; `kernels` implements the per-rank kernel interface."""

    def __init__(self, kernels, group: Optional[dist.ProcessGroup] = None, comm: Optional[NcclComm] = None,
                 peer: Optional[PeerExchange] = None):
        self.k = kernels
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # with a libb200ot-owned communicator the loop is driven from C (one call queues everything)
        self.comm = comm
        # with peer exchange buffers there is no collective call in the loop at all
        self.peer = peer if (peer is not None and hasattr(kernels, "run_peer")) else None
        self.c_loop = self.peer is not None or (comm is not None and hasattr(kernels, "run_c"))
        self.iterations_queued = 0
        self.allreduces = 0
        self._events = []  # bounds how far the host may run ahead of the GPU (see run())
        self._win = max(1, int(os.environ.get("B200OT_SHARD_WINDOW", "10")))
        self._since_event = 0
        # windows the host may run ahead of the device.  With NCCL in the loop 0 was measured best (many outstanding
        # collectives slow NCCL's host side); the peer loop has no library call in it, so the host queues two
        # windows ahead and the device never waits for the next chunk of launches.
        self._lag = max(0, int(os.environ.get("B200OT_SHARD_LAG", "2" if self.peer is not None else "0")))
        self._graph, self._graph_iters = None, 0

    def _allreduce(self, s: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group)
            self.allreduces += 1
        return s

    def start(self):
        """State setup + the first g update (one all-reduce)."""
        self.k.setup()
        if self.peer is not None:
            if self.world > 1 and dist.is_initialized() and torch.cuda.is_available():
                # a new solve reuses the exchange slabs: no rank may still be polling the previous solve's words
                torch.cuda.current_stream().synchronize()
                dist.barrier(group=self.group)
            self.peer.next_epoch()
            self.k.push(self.peer, True)
            self.k.finalize_peer(self.peer, True)
            return
        if self.c_loop:
            self.k.start_c(self.comm)
            self.allreduces += 1
            return
        self.k.finalize(self._allreduce(self.k.prologue()), True)

    def build_graph(self, iters_per_replay: int = 10):
        """Capture `iters_per_replay` iterations (sweep -> reduce -> NCCL all-reduce -> finalize, each) into one
        CUDA graph.  At 8 GPUs an iteration is ~0.4 ms of device time and five enqueues from Python; replaying
        a graph keeps the host out of the loop.  Every kernel starts with `if (state->done) return`, so replaying
        past convergence is harmless.  Runs a short eager warm-up first (kernel attributes, occupancy queries
        and the NCCL communicator must exist before capture); call start() afterwards for the real solve."""
        if not torch.cuda.is_available():
            return None
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.k.setup()
            self.k.finalize(self._allreduce(self.k.prologue()), True)
            for _ in range(2):
                self.k.finalize(self._allreduce(self.k.sweep()), False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(int(iters_per_replay)):
                self.k.finalize(self._allreduce(self.k.sweep()), False)
        self._graph, self._graph_iters = graph, int(iters_per_replay)
        return graph

    def run(self, iters: int):
        """Queue `iters` iterations: local sweep -> all-reduce of m column partials -> finalize.
        Asynchronous apart from a bounded run-ahead window."""
        cuda = torch.cuda.is_available() and self.world > 1
        iters = int(iters)
        if self.c_loop:
            done = 0
            while done < iters:  # chunks keep the host at most two windows ahead of the device
                step = min(self._win, iters - done)
                if self.peer is not None:
                    self.k.run_peer(step, self.peer)
                else:
                    self.k.run_c(step, self.comm)
                done += step
                if torch.cuda.is_available():
                    ev = torch.cuda.Event()
                    ev.record()
                    self._events.append(ev)
                    if len(self._events) > self._lag:
                        self._events.pop(0).synchronize()
            if self.peer is None:
                self.allreduces += iters
            self.iterations_queued += iters
            return
        i = 0
        while i < iters:
            if self._graph is not None and iters - i >= self._graph_iters:
                self._graph.replay()
                step = self._graph_iters
            else:
                self.k.finalize(self._allreduce(self.k.sweep()), False)
                step = 1
            i += step
            self._since_event += step
            # Keep at most 2 windows of iterations queued: with many collectives outstanding NCCL's host side
            # serialises with the device (measured at N=2: 2.7 ms instead of 1.45 ms per iteration when unbounded).
            if cuda and self._since_event >= self._win:
                self._since_event = 0
                ev = torch.cuda.Event()
                ev.record()
                self._events.append(ev)
                if len(self._events) > self._lag:
                    self._events.pop(0).synchronize()
        self.iterations_queued += iters

    def solve(self, max_iter: int, check_every: int = 10, check_phase: int = 1):
        """Blocking solve: chunks that end on check iterations, flags read after each chunk.  A chunk whose fast
        path lost a row or column sum (``bad``: raised on one rank, carried to every rank as NaN column sums, so
        all ranks stop in the same iteration) is rewound to its snapshot and replayed on the robust two-sweep
        kernels -- the same recovery the single-GPU driver ``b200ot_sinkhorn_solve`` performs."""
        self.start()
        done = 0
        ce = max(1, int(check_every))
        robust = False
        can_recover = all(hasattr(self.k, nm) for nm in ("snapshot", "rewind", "use_robust_path"))
        while done < max_iter:
            ln = ((check_phase - done - 1) % ce) + 1
            ln = min(ln, max_iter - done)
            if can_recover:
                self.k.snapshot()
            self.run(ln)
            fl = self.k.flags()
            if fl.get("bad") and can_recover and not robust:
                robust = True
                if self.peer is not None:
                    # the aborted exchanges left words with this epoch's tags in the slabs: new epoch, and no rank
                    # may still be polling when the replay starts to push
                    if self.world > 1 and dist.is_initialized() and torch.cuda.is_available():
                        torch.cuda.current_stream().synchronize()
                        dist.barrier(group=self.group)
                    self.peer.next_epoch()
                self.k.rewind()
                self.k.use_robust_path()
                done = self.k.flags()["it"]
                continue
            done += ln
            if fl["done"]:
                break
        return self.k.finish()


def solve_sharded(C_local, a_local, b, eps, max_iter=1000, tol=1e-9, check_every=10, check_phase=1,
                  err_norm="l2", stop_inclusive=False, path="auto", f0=None, g0=None, group=None, comm=None,
                  peer=None):
    """One call: row-sharded log-domain Sinkhorn; returns this rank's (f_local, g, info).
    Pass a `PeerExchange` (column sums pushed over NVLink as tagged words, no collective call in the loop) or a
    `NcclComm` (ncclAllReduce queued from C on the compute stream); with neither the loop runs from Python."""
    prm = ops.make_params(eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive, path)
    k = CudaShardKernels(C_local, a_local, b, prm, path=path, f0=f0, g0=g0)
    return ShardedSinkhorn(k, group, comm, peer).solve(max_iter, check_every, check_phase)
