"""Online (cost-free) Sinkhorn: C = |x|^2 + |y|^2 - 2 x.y is never materialised in HBM.

The embeddings are split once (two fp16 parts per row by default); every iteration rebuilds C one row panel at a time on the
tensor cores (``b200ot_cost_gemm``) into a panel buffer (default 2 GiB: a few thousand rows -- the whole matrix is
never materialised; ``panel_bytes=48 << 20`` keeps the panel inside the 126 MB L2 instead), and the same
single-sweep kernel that streams a materialised C consumes the panel (``b200ot_sinkhorn_panel_sweep``), adding the
panel's column sums into one vector; ``finalize`` then runs once per iteration exactly as in the streaming solver,
so stopping rule, error history and results are the same.

When to use it: an iteration costs 2*n*m*d*products tensor flops (products = 3 with the
fp16 split, 6 with the bf16 split) instead of 4*n*m bytes of HBM traffic.  At d = 512 that is ~14.7 ms (fp16 split;
24.8 ms with the bf16 split) against ~3 ms per iteration at n = m = 65536, so the streaming path wins whenever C
fits in HBM; the online path is for problems whose cost matrix does not fit (n*m*4 B > ~150 GB).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops
from ._lib import B200OTError, check


class OnlineSinkhorn:
    def __init__(self, x: torch.Tensor, y: torch.Tensor, a: torch.Tensor, b: torch.Tensor, eps: float,
                 max_iter: int = 1000, tol: float = 1e-9, check_every: int = 10, check_phase: int = 1,
                 err_norm: str = "l2", stop_inclusive: bool = False, cost: str = "sqeuclidean", terms=None,
                 panel_bytes: int = 2 << 30, f0: Optional[torch.Tensor] = None, g0: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        x, self.ldx = ops._matrix(x, "x")
        y, self.ldy = ops._matrix(y, "y")
        self.n, self.d = x.shape
        self.m = y.shape[0]
        if y.shape[1] != self.d:
            raise B200OTError("x and y must have the same feature width")
        if self.m % 4:
            raise B200OTError("the online solver needs m % 4 == 0")
        dev = x.device
        self.a = ops._vector(a, "a", self.n)
        self.b = ops._vector(b, "b", self.m)
        self.f0 = None if f0 is None else ops._vector(f0, "f0", self.n)
        self.g0 = None if g0 is None else ops._vector(g0, "g0", self.m)
        self.eps, self.kind = float(eps), _lib.COSTS[cost]
        # terms: C-ABI code of the split; products: tensor products per element (3 with two fp16 parts, 6 with bf16)
        self.terms, self.products, f16 = _lib.split_terms(ops.DEFAULT_COST_TERMS if terms is None else terms)
        self.prm = ops.make_params(eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive, "auto")
        self.max_iter, self.ce, self.cp = int(max_iter), max(1, int(check_every)), int(check_phase)
        # panel height: a multiple of the 128-row GEMM tile.  Measured at n = m = 65536 (bench.py extra.online_c4):
        # L2-sized panels (48 MB = 128 rows) make every iteration 512 GEMM launches of 1.7 waves plus 512 sweeps of
        # 4 rows per cluster -- 49 ms per iteration, 39 % of the bf16 peak; panels of a few thousand rows (2 GiB of
        # HBM scratch, ~1 % of the matrix that is never materialised) give the GEMM full waves and the sweep the
        # same shape as an 8-GPU shard
        rows = max(128, (int(panel_bytes) // (4 * self.m)) // 128 * 128)
        self.panel_rows = min(rows, (self.n + 127) // 128 * 128)
        self.panel = ops.empty_matrix(self.panel_rows, self.m, dev)
        # resident bf16 parts + norms
        def parts(rows_, side):
            nbytes = self.lib.b200ot_cost_parts_bytes(rows_, self.d, side)
            buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            tile = 256 if side else 128
            norms = torch.zeros((rows_ + tile - 1) // tile * tile * (2 if f16 else 1), dtype=torch.float32, device=dev)
            return buf, C.c_void_p((buf.data_ptr() + 1023) // 1024 * 1024), norms
        self._xa, self.xparts, self.xn = parts(self.n, 0)
        self._yb, self.yparts, self.yn = parts(self.m, 1)
        check(self.lib.b200ot_cost_split(ops._ptr(x), self.ldx, self.n, self.d, self.kind, self.terms, 0, self.xparts,
                                         ops._ptr(self.xn), ops._stream()), "b200ot_cost_split(x)")
        check(self.lib.b200ot_cost_split(ops._ptr(y), self.ldy, self.m, self.d, self.kind, self.terms, 1, self.yparts,
                                         ops._ptr(self.yn), ops._stream()), "b200ot_cost_split(y)")
        self.ws = torch.empty(self.lib.b200ot_sinkhorn_workspace_bytes(self.n, self.m) + 256, dtype=torch.uint8,
                              device=dev)
        self.s = torch.zeros(self.m, dtype=torch.float32, device=dev)
        self.tensor_flops_per_iteration = 2.0 * self.n * self.m * self.d * self.products

    def _panels(self):
        for row0 in range(0, self.n, self.panel_rows):
            yield row0, min(self.panel_rows, self.n - row0)

    def _build(self, row0, rows):
        check(self.lib.b200ot_cost_gemm(self.xparts, ops._ptr(self.xn), row0 // 128, rows, self.yparts,
                                        ops._ptr(self.yn), self.m, self.d, self.kind, self.terms,
                                        ops._ptr(self.panel), self.panel.stride(0), ops._stream()), "b200ot_cost_gemm")

    def start(self):
        check(self.lib.b200ot_sinkhorn_setup(self.n, self.m, ops._ptr(self.a), ops._ptr(self.b), ops._ptr(self.f0),
                                             ops._ptr(self.g0), C.byref(self.prm), ops._ws_ptr(self.ws),
                                             self.ws.numel() - 256, ops._stream()), "b200ot_sinkhorn_setup")
        for row0, rows in self._panels():
            self._build(row0, rows)
            check(self.lib.b200ot_sinkhorn_panel_prologue(ops._ptr(self.panel), self.panel.stride(0), self.n, self.m,
                                                          row0, rows, ops._ws_ptr(self.ws), ops._ptr(self.s),
                                                          int(row0 > 0), ops._stream()), "b200ot_sinkhorn_panel_prologue")
        check(self.lib.b200ot_sinkhorn_shard_finalize(self.n, self.m, ops._ws_ptr(self.ws), ops._ptr(self.s), 1,
                                                      ops._stream()), "b200ot_sinkhorn_shard_finalize")

    def run(self, iters: int):
        for _ in range(int(iters)):
            for row0, rows in self._panels():
                self._build(row0, rows)
                check(self.lib.b200ot_sinkhorn_panel_sweep(ops._ptr(self.panel), self.panel.stride(0), self.n, self.m,
                                                           row0, rows, _lib.PATH_AUTO, ops._ws_ptr(self.ws),
                                                           ops._ptr(self.s), int(row0 > 0), ops._stream()),
                      "b200ot_sinkhorn_panel_sweep")
            check(self.lib.b200ot_sinkhorn_shard_finalize(self.n, self.m, ops._ws_ptr(self.ws), ops._ptr(self.s), 0,
                                                          ops._stream()), "b200ot_sinkhorn_shard_finalize")

    def flags(self):
        out = torch.empty(8, dtype=torch.int32, device=self.s.device)
        check(self.lib.b200ot_sinkhorn_peek(ops._ws_ptr(self.ws), ops._ptr(out), ops._stream()), "b200ot_sinkhorn_peek")
        v = out.cpu().tolist()
        return {"it": v[0], "done": v[1], "converged": v[2], "bad": v[4], "n_err": v[5]}

    def finish(self):
        f = torch.empty(self.n, dtype=torch.float32, device=self.s.device)
        g = torch.empty(self.m, dtype=torch.float32, device=self.s.device)
        res = torch.zeros(8, dtype=torch.int32, device=self.s.device)
        errs = torch.zeros(512, dtype=torch.float32, device=self.s.device)
        check(self.lib.b200ot_sinkhorn_finish(self.n, self.m, ops._ws_ptr(self.ws), ops._ptr(f), ops._ptr(g),
                                              ops._ptr(res), ops._ptr(errs), 512, ops._stream()), "b200ot_sinkhorn_finish")
        r = res.cpu()
        n_err = int(r[3])
        return f, g, {"n_iter": int(r[0]), "converged": bool(r[1]), "status": int(r[2]), "n_err": n_err,
                      "err": float(r[4:5].view(torch.float32)[0]), "errs": errs[:min(n_err, 512)]}

    def solve(self):
        self.start()
        done = 0
        while done < self.max_iter:
            ln = min(((self.cp - done - 1) % self.ce) + 1, self.max_iter - done)
            self.run(ln)
            done += ln
            if self.flags()["done"]:
                break
        return self.finish()


# Rates measured on one B200 (round 2; bench.py `roofline` / `extra.online_c4`, profiles/r02_bench_n1.json,
# profiles/r02_online_launches.csv).  choose_path compares the two per-iteration costs with THESE numbers, not
# with data-sheet peaks.
MEASURED = {
    "stream_hbm_gbs": 5900.0,           # single-sweep kernel at 65536^2: 0.85-0.96 of the 6537 GB/s copy peak across boxes
    "online_tflops_executed": 890.0,    # cost_tc panels of 8192 rows + panel sweep, fp16 split: 3 products per element
    "online_products": 3,
    "online_min_k": 64,                 # the GEMM pads d to its 64-wide K block
}


def choose_path(n: int, m: int, d: int, free_bytes: int, terms: Optional[int] = None, rates: Optional[dict] = None) -> str:
    """'streaming' (materialise C once, 4nm bytes of HBM per iteration) or 'online' (2nm*d*products tensor flops per
    iteration plus one write and one read of every panel, no n x m matrix).  Counter-driven: both per-iteration
    times are evaluated with the measured rates in ``MEASURED``.  At d = 512 streaming is ~5x faster (fp16 split),
    and because every panel is written once and read once the online path never beats a resident C even for
    narrow embeddings: it is chosen when C and its workspace do not fit in HBM.  `terms`: tensor products per
    element if not the default split's."""
    r = dict(MEASURED if rates is None else rates)
    nm = float(n) * float(m)
    t_stream = 4.0 * nm / (r["stream_hbm_gbs"] * 1e9)
    d_eff = max(int(d), int(r.get("online_min_k", 64)))
    products = int(r.get("online_products", 3)) if terms is None else int(terms)
    t_online = 2.0 * nm * d_eff * products / (r["online_tflops_executed"] * 1e12) + 8.0 * nm / (r["stream_hbm_gbs"] * 1e9)
    fits = 4.0 * nm * 1.02 + 3e8 <= free_bytes
    if fits and t_stream <= t_online:
        return "streaming"
    return "online"
