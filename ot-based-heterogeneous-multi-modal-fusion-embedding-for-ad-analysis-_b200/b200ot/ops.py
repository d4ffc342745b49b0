"""Device-level operators: thin, typed wrappers of the C ABI on CUDA torch tensors.

PyTorch is used for device memory and streams only; every computation below is a
call into libb200ot.so.  Nothing here falls back to torch math or to the CPU.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import torch

from . import _lib
from ._lib import B200OTError, Params, Result, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_tensors(obj):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            yield from _cuda_tensors(o)


def on_device(fn):
    """Run an operator on the device its tensors live on: every CUDA operand must share ONE device, and the call
    (stream lookup, kernel launches, per-device function attributes inside the library) happens with that device
    current -- tensors on cuda:1 while cuda:0 is current would otherwise be launched on the wrong device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for t in _cuda_tensors(list(args) + list(kwargs.values())):
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise B200OTError(f"{fn.__name__}: operands live on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B200OTError(f"{name} must be a CUDA tensor (b200ot has no CPU path)")
    if t.dtype != dtype:
        raise B200OTError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _matrix(t: torch.Tensor, name: str):
    """fp32 CUDA matrix with unit column stride; returns (tensor, ld)."""
    _need_cuda(t, name)
    if t.dim() != 2:
        raise B200OTError(f"{name} must be 2-D")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t, t.stride(0)


def _vector(t: torch.Tensor, name: str, length: int, dtype=torch.float32):
    _need_cuda(t, name, dtype)
    t = t.reshape(-1)
    if t.numel() != length:
        raise B200OTError(f"{name} has {t.numel()} entries, expected {length}")
    return t.contiguous()


def empty_matrix(n: int, m: int, device) -> torch.Tensor:
    """n x m fp32 matrix whose row stride is a multiple of 4 floats and whose base is
    16-byte aligned, i.e. eligible for the single-sweep Sinkhorn kernel."""
    ld = (m + 3) // 4 * 4
    return torch.empty((n, ld), dtype=torch.float32, device=device)[:, :m]


def aligned_copy(Cm: torch.Tensor) -> torch.Tensor:
    """Return Cm itself if it already meets the fused kernel's layout, else a padded copy."""
    if (Cm.stride(1) == 1 and Cm.stride(0) % 4 == 0 and Cm.data_ptr() % 16 == 0
            and Cm.shape[1] % 4 == 0):
        return Cm
    if Cm.shape[1] % 4 != 0:
        return Cm  # generic kernels handle it; padding cannot make m a multiple of 4
    out = empty_matrix(Cm.shape[0], Cm.shape[1], Cm.device)
    out.copy_(Cm)
    return out


# ---------------------------------------------------------------------------
# cost construction
# ---------------------------------------------------------------------------
# two fp16 parts per operand, 3 products: fp32-grade relative to |x||y| at half the tensor work of the 6-term bf16
# split (tests/test_gpu_parity.py::test_cost_fp16_split_scales_every_row)
DEFAULT_COST_TERMS = "f16"
@on_device
def cost_matrix(x: torch.Tensor, y: torch.Tensor, kind: str = "sqeuclidean",
                out: Optional[torch.Tensor] = None, impl: str = "auto", terms=DEFAULT_COST_TERMS) -> torch.Tensor:
    """C_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  or  1 - cos(x_i, y_j).

    impl="tc": tcgen05 split GEMM (b200ot_cost); impl="simt": fp32 FMA kernel (b200ot_cost_simt); "auto" picks the
    tensor-core kernel once the problem fills the GPU.  terms: "f16" = rows scaled by a power of two and split into
    two fp16 parts, 3 products (fp32-grade relative to |x||y|, half the tensor work of 6); 6 / 3 / 1 = bf16 parts."""
    lib = _lib.load()
    x, ldx = _matrix(x, "x")
    y, ldy = _matrix(y, "y")
    n, d = x.shape
    m, d2 = y.shape
    if d != d2:
        raise B200OTError("x and y must have the same feature width")
    if out is None:
        out = empty_matrix(n, m, x.device)
    out, ldc = _matrix(out, "out")
    if impl == "auto":
        impl = "tc" if (n >= 512 and m >= 512 and d >= 64) else "simt"
    if impl == "tc":
        need = lib.b200ot_cost_workspace_bytes(n, m, d)
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
        wsp = C.c_void_p((ws.data_ptr() + 1023) // 1024 * 1024)
        check(lib.b200ot_cost(_ptr(x), ldx, _ptr(y), ldy, n, m, d, _lib.COSTS[kind], _ptr(out), ldc, wsp,
                              need, _lib.split_terms(terms)[0], _stream()), "b200ot_cost")
    elif impl == "simt":
        norms = torch.empty(n + m, dtype=torch.float32, device=x.device)
        check(lib.b200ot_cost_simt(_ptr(x), ldx, _ptr(y), ldy, n, m, d, _lib.COSTS[kind], _ptr(out), ldc,
                                   _ptr(norms), _stream()), "b200ot_cost_simt")
    else:
        raise B200OTError(f"unknown cost impl {impl!r}")
    return out


@on_device
def fot_cost(A: torch.Tensor, B: torch.Tensor, Ts: torch.Tensor, w1: torch.Tensor,
             w2: torch.Tensor, impl: str = "auto") -> torch.Tensor:
    """M = (A.^2)^T w1 (+) (B.^2)^T w2 - 2 A^T Ts B.  impl = "tc": both contractions on tcgen05
    (b200ot_fot_cost_tc), "simt": fp32 FMA (b200ot_fot_cost), "auto": tensor cores from n d d' >= 2^29 (128 samples
    of 2048 features, where the two are equal at 0.12 ms: the chain is ten small launches; 4.6x at 2048 x 4096)."""
    lib = _lib.load()
    A, lda = _matrix(A, "A")
    B, ldb = _matrix(B, "B")
    Ts, ldt = _matrix(Ts, "Ts")
    n, d = A.shape
    n2, d2 = B.shape
    if Ts.shape != (n, n2):
        raise B200OTError(f"Ts must be {n} x {n2}")
    if impl not in ("auto", "tc", "simt"):
        raise B200OTError(f"impl must be auto, tc or simt, got {impl!r}")
    w1 = _vector(w1, "w1", n)
    w2 = _vector(w2, "w2", n2)
    M = empty_matrix(d, d2, A.device)
    if impl == "tc" or (impl == "auto" and n * d * d2 >= (1 << 29)):
        nbytes = lib.b200ot_fot_cost_tc_workspace_bytes(n, n2, d, d2)
        buf, ws = _tc_ws(nbytes, A.device)
        check(lib.b200ot_fot_cost_tc(_ptr(A), lda, _ptr(B), ldb, _ptr(Ts), ldt, _ptr(w1), _ptr(w2), n, n2, d, d2,
                                     _ptr(M), M.stride(0), ws, nbytes, _stream()), "b200ot_fot_cost_tc")
        return M
    tmp = torch.empty(n * d2 + d + d2, dtype=torch.float32, device=A.device)
    check(lib.b200ot_fot_cost(_ptr(A), lda, _ptr(B), ldb, _ptr(Ts), ldt, _ptr(w1), _ptr(w2), n, n2, d, d2,
                              _ptr(M), M.stride(0), _ptr(tmp), _stream()), "b200ot_fot_cost")
    return M


@on_device
def matrix_max(Cm: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    out = torch.empty(1, dtype=torch.float32, device=Cm.device)
    check(lib.b200ot_matrix_max(_ptr(Cm), ldc, Cm.shape[0], Cm.shape[1], _ptr(out), _stream()),
          "b200ot_matrix_max")
    return out


@on_device
def scale_by_inv_(Cm: torch.Tensor, denom: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    if Cm.stride(1) != 1:
        raise B200OTError("in-place scale needs unit column stride")
    check(lib.b200ot_matrix_scale_by_inv(_ptr(Cm), Cm.stride(0), Cm.shape[0], Cm.shape[1], _ptr(denom),
                                         _stream()), "b200ot_matrix_scale_by_inv")
    return Cm


# ---------------------------------------------------------------------------
# Sinkhorn
# ---------------------------------------------------------------------------
_WS_CACHE: dict = {}


def _workspace(n: int, m: int, device) -> torch.Tensor:
    lib = _lib.load()
    need = lib.b200ot_sinkhorn_workspace_bytes(n, m)
    key = (torch.device(device).index, n, m)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < need:
        if len(_WS_CACHE) > 8:
            _WS_CACHE.clear()
        ws = torch.empty(need + 256, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def _ws_ptr(ws: torch.Tensor):
    base = ws.data_ptr()
    return C.c_void_p((base + 255) // 256 * 256)


def describe_kernel(n: int, m: int) -> str:
    """Which Sinkhorn kernel configuration the library picks for an n x m problem (needs a GPU)."""
    lib = _lib.load()
    buf = C.create_string_buffer(512)
    check(lib.b200ot_sinkhorn_describe(int(n), int(m), buf, 512), "b200ot_sinkhorn_describe")
    return buf.value.decode()


def make_params(eps, max_iter, tol, check_every=10, check_phase=1, err_norm="l2",
                stop_inclusive=False, path="auto", floor_patience=0) -> Params:
    return Params(float(eps), int(max_iter), float(tol), int(check_every), int(check_phase),
                  _lib.NORMS[err_norm], int(bool(stop_inclusive)), _lib.PATHS[path], int(floor_patience))


@on_device
def sinkhorn_potentials(Cm: torch.Tensor, a: torch.Tensor, b: torch.Tensor, eps: float,
                        max_iter: int = 1000, tol: float = 1e-9, check_every: int = 10,
                        check_phase: int = 1, err_norm: str = "l2", stop_inclusive: bool = False,
                        path: str = "auto", f0: Optional[torch.Tensor] = None,
                        g0: Optional[torch.Tensor] = None, err_hist_cap: int = 512, floor_patience: int = 0):
    """Log-domain Sinkhorn on a cost matrix resident in HBM (b200ot_sinkhorn_solve).

    Returns ``(f, g, info)`` with ``P = exp((f_i + g_j - C_ij)/eps)``; ``info`` has
    ``n_iter``, ``converged``, ``err``, ``errs`` (error history), ``status``, ``time``.
    """
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    a = _vector(a, "a", n)
    b = _vector(b, "b", m)
    f0 = None if f0 is None else _vector(f0, "f0", n)
    g0 = None if g0 is None else _vector(g0, "g0", m)
    ws = _workspace(n, m, Cm.device)
    f = torch.empty(n, dtype=torch.float32, device=Cm.device)
    g = torch.empty(m, dtype=torch.float32, device=Cm.device)
    errs = torch.zeros(max(1, err_hist_cap), dtype=torch.float32, device=Cm.device)
    prm = make_params(eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive, path,
                      floor_patience)
    res = Result()
    t0 = time.perf_counter()
    check(lib.b200ot_sinkhorn_solve(_ptr(Cm), ldc, n, m, _ptr(a), _ptr(b), _ptr(f0), _ptr(g0),
                                    C.byref(prm), _ws_ptr(ws), ws.numel() - 256, _ptr(f), _ptr(g),
                                    C.byref(res), _ptr(errs), errs.numel(), _stream()),
          "b200ot_sinkhorn_solve")
    info = {"n_iter": res.n_iter, "converged": bool(res.converged), "err": res.err,
            "status": res.status, "n_err": res.n_err, "time": time.perf_counter() - t0,
            "errs": errs[:min(res.n_err, errs.numel())]}
    return f, g, info


def _on_self_device(method):
    """Method form of on_device for objects that own their matrix (``self.C``)."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        dev = self.C.device
        if dev.index == torch.cuda.current_device():
            return method(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return method(self, *args, **kwargs)

    return wrapper


class SinkhornStepper:
    """Asynchronous life cycle (init / enqueue / finish) for callers that drive the
    iteration themselves: bench loops, CUDA-graph capture, the row-sharded solver."""

    def __init__(self, Cm: torch.Tensor, a, b, eps, max_iter=1000, tol=0.0, check_every=10,
                 check_phase=1, err_norm="l2", stop_inclusive=False, path="auto", f0=None, g0=None):
        self.lib = _lib.load()
        self.C, self.ldc = _matrix(Cm, "C")
        self.n, self.m = self.C.shape
        self.a = _vector(a, "a", self.n)
        self.b = _vector(b, "b", self.m)
        self.prm = make_params(eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive,
                               path)
        self.path = _lib.PATHS[path]
        self.ws = torch.empty(self.lib.b200ot_sinkhorn_workspace_bytes(self.n, self.m) + 256,
                              dtype=torch.uint8, device=self.C.device)
        self.f0 = None if f0 is None else _vector(f0, "f0", self.n)
        self.g0 = None if g0 is None else _vector(g0, "g0", self.m)
        self.reset()

    @_on_self_device
    def reset(self):
        check(self.lib.b200ot_sinkhorn_init(_ptr(self.C), self.ldc, self.n, self.m, _ptr(self.a),
                                            _ptr(self.b), _ptr(self.f0), _ptr(self.g0),
                                            C.byref(self.prm), _ws_ptr(self.ws), self.ws.numel() - 256,
                                            _stream()), "b200ot_sinkhorn_init")

    @_on_self_device
    def enqueue(self, iters: int):
        check(self.lib.b200ot_sinkhorn_enqueue(_ptr(self.C), self.ldc, self.n, self.m, int(iters),
                                               self.path, _ws_ptr(self.ws), _stream()),
              "b200ot_sinkhorn_enqueue")

    @_on_self_device
    def build_graph(self, iters_per_replay: int = 10):
        """Capture `iters_per_replay` iterations into a CUDA graph (launch-bound small problems: two kernel
        launches per iteration become one graph replay per `iters_per_replay`).  Kernels no-op once the
        stopping rule has fired, so replaying past convergence is harmless."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.reset()
            self.enqueue(2)  # warm-up: kernel attributes and occupancy queries happen outside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.enqueue(int(iters_per_replay))
        self._graph_iters = int(iters_per_replay)
        self.reset()
        return self._graph

    @_on_self_device
    def run(self, iters: int):
        """enqueue() through the captured graph where whole replays fit, eagerly for the remainder."""
        g = getattr(self, "_graph", None)
        iters = int(iters)
        while g is not None and iters >= self._graph_iters:
            g.replay()
            iters -= self._graph_iters
        if iters > 0:
            self.enqueue(iters)

    @_on_self_device
    def flags(self) -> dict:
        out = torch.empty(8, dtype=torch.int32, device=self.C.device)
        check(self.lib.b200ot_sinkhorn_peek(_ws_ptr(self.ws), _ptr(out), _stream()), "b200ot_sinkhorn_peek")
        v = out.cpu().tolist()
        return {"it": v[0], "done": v[1], "converged": v[2], "cur": v[3], "bad": v[4], "n_err": v[5]}

    @_on_self_device
    def finish(self, err_hist_cap: int = 512):
        f = torch.empty(self.n, dtype=torch.float32, device=self.C.device)
        g = torch.empty(self.m, dtype=torch.float32, device=self.C.device)
        res = torch.zeros(8, dtype=torch.int32, device=self.C.device)
        errs = torch.zeros(err_hist_cap, dtype=torch.float32, device=self.C.device)
        check(self.lib.b200ot_sinkhorn_finish(self.n, self.m, _ws_ptr(self.ws), _ptr(f), _ptr(g), _ptr(res),
                                              _ptr(errs), err_hist_cap, _stream()), "b200ot_sinkhorn_finish")
        r = res.cpu()
        n_err = int(r[3])
        info = {"n_iter": int(r[0]), "converged": bool(r[1]), "status": int(r[2]), "n_err": n_err,
                "err": float(r[4:5].view(torch.float32)[0]), "errs": errs[:min(n_err, err_hist_cap)]}
        return f, g, info


@on_device
def sinkhorn_batched(a: torch.Tensor, b: torch.Tensor, eps: float, *, C3: Optional[torch.Tensor] = None,
                     X: Optional[torch.Tensor] = None, Y: Optional[torch.Tensor] = None,
                     max_iter: int = 1000, tol: float = 1e-9, check_every: int = 10, check_phase: int = 1,
                     err_norm: str = "l2", stop_inclusive: bool = False):
    """Batch of independent small problems, one per CTA, float64 kernel-domain Sinkhorn-Knopp
    (b200ot_sinkhorn_batched).  Give either C3 (B x n x m) or embeddings X (B x n x d), Y (B x m x d)."""
    lib = _lib.load()
    if C3 is not None:
        _need_cuda(C3, "C3")
        C3 = C3.contiguous()
        B, n, m = C3.shape
        d = 0
        dev = C3.device
    else:
        _need_cuda(X, "X")
        _need_cuda(Y, "Y")
        X = X.contiguous()
        Y = Y.contiguous()
        B, n, d = X.shape
        B2, m, d2 = Y.shape
        if B != B2 or d != d2:
            raise B200OTError("X and Y batch / width mismatch")
        dev = X.device
    a = _vector(a, "a", n)
    b = _vector(b, "b", m)
    P = torch.empty((B, n, m), dtype=torch.float32, device=dev)
    u = torch.empty((B, n), dtype=torch.float64, device=dev)
    v = torch.empty((B, m), dtype=torch.float64, device=dev)
    n_iter = torch.empty(B, dtype=torch.int32, device=dev)
    err = torch.empty(B, dtype=torch.float32, device=dev)
    prm = make_params(eps, max_iter, tol, check_every, check_phase, err_norm, stop_inclusive)
    check(lib.b200ot_sinkhorn_batched(_ptr(C3), _ptr(X), _ptr(Y), B, n, m, d, _ptr(a), _ptr(b),
                                      C.byref(prm), _ptr(P), _ptr(u), _ptr(v), _ptr(n_iter), _ptr(err),
                                      _stream()), "b200ot_sinkhorn_batched")
    return P, {"u": u, "v": v, "n_iter": n_iter, "err": err}


# ---------------------------------------------------------------------------
# epilogues
# ---------------------------------------------------------------------------
@on_device
def plan(Cm, f, g, eps, out=None):
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    f = _vector(f, "f", n)
    g = _vector(g, "g", m)
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=Cm.device)
    out, ldp = _matrix(out, "out")
    check(lib.b200ot_plan(_ptr(Cm), ldc, n, m, _ptr(f), _ptr(g), float(eps), _ptr(out), ldp, _stream()),
          "b200ot_plan")
    return out


@on_device
def ot_cost(Cm, f, g, eps):
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    out = torch.empty(_lib.OT_COST_DOUBLES, dtype=torch.float64, device=Cm.device)
    check(lib.b200ot_ot_cost(_ptr(Cm), ldc, n, m, _ptr(_vector(f, "f", n)), _ptr(_vector(g, "g", m)),
                             float(eps), _ptr(out), _stream()), "b200ot_ot_cost")
    return out[:1]


@on_device
def plan_guard_rownorm(Cm=None, f=None, g=None, eps=None, T=None, out=None):
    """The per-step plan guard of the reference (MRI_PET_OT_nojax.py:704-715) in one kernel: NaN -> 1e-8, rows
    divided by their sums (0 -> 1e-8).  Either ``(Cm, f, g, eps)`` (plan evaluated on the fly) or a dense ``T``."""
    lib = _lib.load()
    if T is not None:
        T, ldt = _matrix(T, "T")
        n, m = T.shape
        if out is None:
            out = torch.empty((n, m), dtype=torch.float32, device=T.device)
        out, ldp = _matrix(out, "out")
        check(lib.b200ot_plan_guard_rownorm(None, 0, n, m, None, None, 1.0, _ptr(T), ldt, _ptr(out), ldp, _stream()),
              "b200ot_plan_guard_rownorm")
        return out
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=Cm.device)
    out, ldp = _matrix(out, "out")
    check(lib.b200ot_plan_guard_rownorm(_ptr(Cm), ldc, n, m, _ptr(_vector(f, "f", n)), _ptr(_vector(g, "g", m)),
                                        float(eps), None, 0, _ptr(out), ldp, _stream()), "b200ot_plan_guard_rownorm")
    return out


_TC_MAX_DV = 512


def _tc_ws(nbytes: int, device) -> tuple:
    """1024-byte aligned scratch of `nbytes` for the tensor-core epilogues (torch owns the memory)."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return buf, C.c_void_p((buf.data_ptr() + 1023) // 1024 * 1024)


def _apply_wants_tc(n: int, m: int, dv: int) -> bool:
    # the tcgen05 kernel pads dv to 16 and works on 128-row blocks: tiny problems stay on the SIMT kernel
    return n * m >= (1 << 18) and dv >= 16


@on_device
def apply_plan(Cm, f, g, eps, V, normalise=False, transpose=False, impl="auto", return_rowsum=False, weights=None):
    """Z = P V (or P^T V with transpose=True), optionally row-normalised, without forming P.

    impl="tc": single-pass tcgen05 kernel (b200ot_apply_plan_tc; dv is processed in slabs of 512 columns);
    impl="simt": the generic fp32 kernel (any shape / alignment); "auto" picks by size.  With
    ``return_rowsum`` the row sums of P (P^T when transposed) come back as well.

    ``weights=(w0, w1, wrow, wcol)``: contract with ``W_ij = P_ij (w0 + w1 C_ij + wrow_i + wcol_j)`` instead of P
    (tensor-core kernel only; ``wrow`` over the rows of C, ``wcol`` over its columns, either may be None): the
    C-weighted products and the implicit-gradient weights of ``torch_ops.ot_loss(grad="implicit")``."""
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    V, ldv = _matrix(V, "V")
    rows_in, rows_out = (n, m) if transpose else (m, n)
    if V.shape[0] != rows_in:
        raise B200OTError(f"V must have {rows_in} rows")
    dv = V.shape[1]
    fv, gv = _vector(f, "f", n), _vector(g, "g", m)
    Z = torch.empty((rows_out, dv), dtype=torch.float32, device=Cm.device)
    if impl == "auto":
        impl = "tc" if (_apply_wants_tc(n, m, dv) or weights is not None) else "simt"
    if weights is not None:
        if impl != "tc" or normalise:
            raise B200OTError("weighted plan products run on the tensor-core kernel, without normalisation")
        w0, w1, wrow, wcol = weights
        wrow = None if wrow is None else _vector(wrow, "wrow", n)
        wcol = None if wcol is None else _vector(wcol, "wcol", m)
        rs = torch.empty(rows_out, dtype=torch.float32, device=Cm.device) if return_rowsum else None
        for c0 in range(0, dv, _TC_MAX_DV):
            dvc = min(_TC_MAX_DV, dv - c0)
            need = lib.b200ot_apply_plan_tc_workspace_bytes(n, m, dvc, int(bool(transpose)))
            buf, wsp = _tc_ws(need, Cm.device)
            Vc, Zc = V[:, c0:c0 + dvc], Z[:, c0:c0 + dvc]
            check(lib.b200ot_apply_plan_tc_weighted(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), float(w0),
                                                    float(w1), _ptr(wrow), _ptr(wcol), _ptr(Vc), ldv, dvc,
                                                    int(bool(transpose)), _ptr(Zc), Z.stride(0),
                                                    _ptr(rs) if c0 == 0 else None, wsp, need, _stream()),
                  "b200ot_apply_plan_tc_weighted")
        return (Z, rs) if return_rowsum else Z
    if impl == "tc":
        rs = torch.empty(rows_out, dtype=torch.float32, device=Cm.device) if return_rowsum else None
        for c0 in range(0, dv, _TC_MAX_DV):
            dvc = min(_TC_MAX_DV, dv - c0)
            need = lib.b200ot_apply_plan_tc_workspace_bytes(n, m, dvc, int(bool(transpose)))
            buf, wsp = _tc_ws(need, Cm.device)
            Vc, Zc = V[:, c0:c0 + dvc], Z[:, c0:c0 + dvc]
            check(lib.b200ot_apply_plan_tc(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), _ptr(Vc), ldv, dvc,
                                           int(bool(transpose)), int(bool(normalise)), _ptr(Zc), Z.stride(0),
                                           _ptr(rs) if c0 == 0 else None, wsp, need, _stream()),
                  "b200ot_apply_plan_tc")
        return (Z, rs) if return_rowsum else Z
    if impl != "simt":
        raise B200OTError(f"unknown apply_plan impl {impl!r}")
    fn = lib.b200ot_apply_plan_t if transpose else lib.b200ot_apply_plan
    check(fn(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), _ptr(V),
             ldv, dv, int(bool(normalise)), _ptr(Z), Z.stride(0), _stream()), "b200ot_apply_plan")
    if return_rowsum:
        ones = torch.ones((rows_in, 1), dtype=torch.float32, device=Cm.device)
        rs = torch.empty((rows_out, 1), dtype=torch.float32, device=Cm.device)
        check(fn(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), _ptr(ones), 1, 1, 0, _ptr(rs), 1, _stream()),
              "b200ot_apply_plan")
        return Z, rs.reshape(-1)
    return Z


@on_device
def envelope_bwd(Cm, f, g, eps, x, y, scale: float = 2.0, impl="auto", weights=None):
    """dX = scale (diag(P1) x - P y), dY = scale (diag(P^T 1) y - P^T x): the gradient of <P, C(x, y)> for the
    squared-Euclidean cost with the plan held fixed.  impl="tc": both halves in ONE launch of the tcgen05 kernel
    (b200ot_envelope_bwd); "simt": two plan-free products on the generic kernel."""
    lib = _lib.load()
    Cm, ldc = _matrix(Cm, "C")
    n, m = Cm.shape
    x, ldx = _matrix(x, "x")
    y, ldy = _matrix(y, "y")
    d = x.shape[1]
    if x.shape[0] != n or y.shape[0] != m or y.shape[1] != d:
        raise B200OTError("envelope_bwd: x must be n x d and y m x d")
    fv, gv = _vector(f, "f", n), _vector(g, "g", m)
    if impl == "auto":
        impl = "tc" if ((_apply_wants_tc(n, m, d) or weights is not None) and d <= _TC_MAX_DV) else "simt"
    if weights is not None and impl != "tc":
        raise B200OTError("the weighted (implicit) gradient runs on the tensor-core kernel (d <= 512)")
    if impl == "tc":
        dx = torch.empty((n, d), dtype=torch.float32, device=Cm.device)
        dy = torch.empty((m, d), dtype=torch.float32, device=Cm.device)
        need = lib.b200ot_envelope_bwd_workspace_bytes(n, m, d)
        buf, wsp = _tc_ws(need, Cm.device)
        if weights is not None:
            w0, w1, wrow, wcol = weights
            wrow = None if wrow is None else _vector(wrow, "wrow", n)
            wcol = None if wcol is None else _vector(wcol, "wcol", m)
            check(lib.b200ot_envelope_bwd_weighted(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), float(w0),
                                                   float(w1), _ptr(wrow), _ptr(wcol), _ptr(x), ldx, _ptr(y), ldy, d,
                                                   float(scale), _ptr(dx), d, _ptr(dy), d, wsp, need, _stream()),
                  "b200ot_envelope_bwd_weighted")
            return dx, dy
        check(lib.b200ot_envelope_bwd(_ptr(Cm), ldc, n, m, _ptr(fv), _ptr(gv), float(eps), _ptr(x), ldx, _ptr(y), ldy,
                                      d, float(scale), _ptr(dx), d, _ptr(dy), d, None, None, wsp, need, _stream()),
              "b200ot_envelope_bwd")
        return dx, dy
    Py, r = apply_plan(Cm, fv, gv, eps, y, impl="simt", return_rowsum=True)
    Ptx, c = apply_plan(Cm, fv, gv, eps, x, transpose=True, impl="simt", return_rowsum=True)
    return scale * (r[:, None] * x - Py), scale * (c[:, None] * y - Ptx)


@on_device
def cosine_loss(A, B):
    lib = _lib.load()
    A, lda = _matrix(A, "A")
    B, ldb = _matrix(B, "B")
    if A.shape != B.shape:
        raise B200OTError("cosine_loss operands must have the same shape")
    out = torch.empty(1, dtype=torch.float32, device=A.device)
    check(lib.b200ot_cosine_loss(_ptr(A), lda, _ptr(B), ldb, A.shape[0], A.shape[1], _ptr(out), _stream()),
          "b200ot_cosine_loss")
    return out


@on_device
def foscttm(pred: torch.Tensor, true: torch.Tensor) -> torch.Tensor:
    """Fraction of samples closer than the true match, per sample (b200ot_cost + b200ot_foscttm)."""
    lib = _lib.load()
    if pred.shape != true.shape:
        raise B200OTError("foscttm operands must have the same shape")
    D = cost_matrix(pred, true, impl="simt" if pred.shape[0] < 512 else "auto")
    out = torch.empty(pred.shape[0], dtype=torch.float32, device=pred.device)
    check(lib.b200ot_foscttm(_ptr(D), D.stride(0), pred.shape[0], _ptr(out), _stream()), "b200ot_foscttm")
    return out


@on_device
def egw_batched(Xs, Ys, eps: float = 5e-3, gw_max_iter: int = 2000, sk_max_iter: int = 2000, gw_threshold: float = 1e-3,
                gw_min_iter: int = 5, sk_threshold: float = 1e-3, sk_check_every: int = 10):
    """Entropic Gromov-Wasserstein couplings for a list of independent small problems (one per label), one CTA
    each (b200ot_egw_batched).  Xs[l]: (n_l, dx) CUDA fp32, Ys[l]: (m_l, dy); n_l, m_l <= 64.
    Returns ``(Ts, info)``: float64 couplings and a dict of per-problem tensors."""
    lib = _lib.load()
    if len(Xs) == 0 or len(Xs) != len(Ys):
        raise B200OTError("egw_batched needs matching, non-empty lists of point clouds")
    dev = Xs[0].device
    for t in list(Xs) + list(Ys):
        _need_cuda(t, "point cloud")
        if t.dim() != 2 or t.shape[0] < 1:
            raise B200OTError("point clouds must be non-empty 2-D tensors")
    dx, dy = Xs[0].shape[1], Ys[0].shape[1]
    if any(x.shape[1] != dx for x in Xs) or any(y.shape[1] != dy for y in Ys):
        raise B200OTError("all point clouds of one side must share their width")
    ns = [int(x.shape[0]) for x in Xs]
    ms = [int(y.shape[0]) for y in Ys]
    X = torch.cat([x.contiguous() for x in Xs]).contiguous()
    Y = torch.cat([y.contiguous() for y in Ys]).contiguous()
    xoff = torch.tensor([0] + list(torch.tensor(ns).cumsum(0).tolist()), dtype=torch.int32, device=dev)
    yoff = torch.tensor([0] + list(torch.tensor(ms).cumsum(0).tolist()), dtype=torch.int32, device=dev)
    sizes = [a * b for a, b in zip(ns, ms)]
    toff_h = [0]
    for sz in sizes:
        toff_h.append(toff_h[-1] + sz)
    toff = torch.tensor(toff_h, dtype=torch.int64, device=dev)
    T = torch.empty(toff_h[-1], dtype=torch.float64, device=dev)
    info = torch.zeros((len(ns), 4), dtype=torch.int32, device=dev)
    cost = torch.zeros(len(ns), dtype=torch.float64, device=dev)
    check(lib.b200ot_egw_batched(_ptr(X), _ptr(Y), _ptr(xoff), _ptr(yoff), _ptr(toff), len(ns), max(ns), max(ms),
                                 dx, dy, float(eps), int(gw_max_iter), int(gw_min_iter), float(gw_threshold),
                                 int(sk_max_iter), int(sk_check_every), float(sk_threshold), _ptr(T), _ptr(info),
                                 _ptr(cost), _stream()), "b200ot_egw_batched")
    Ts = [T[toff_h[i]:toff_h[i + 1]].view(ns[i], ms[i]) for i in range(len(ns))]
    return Ts, {"n_iters_outer": info[:, 0], "converged_outer": info[:, 1], "converged_inner": info[:, 2],
                "inner_iterations": info[:, 3], "GW cost": cost}
