// Solver state shared by the Sinkhorn kernels (sinkhorn.cu: one launch per sweep; resident.cu: whole
// solves in one persistent launch): the 256-byte state block, the workspace layout and the state machine
// that applies the reference's stopping rule (perturbot/perturbot/match/utils.py:48,80-89; POT / ott
// flavours per SURVEY appendix A).
#pragma once
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b200ot {

constexpr int kErrHistCap = 4096;
constexpr int kNpCap = 160;  // max partial slabs (>= clusters of the fused sweep, row splits of the robust one)
constexpr int kFinalizeThreads = 256;
constexpr int kResMaxCtas = 320;  // resident kernel: 2 CTAs per SM
constexpr int kResMaxM = 8192;    // widest row one resident CTA covers

struct State {
  int it, done, converged, cur;
  int bad, n_err, ticket, initialised;
  float err, kscale, eps, tol;
  int max_iter, check_every, check_phase, err_norm;
  int stop_inclusive, path, snap_it, snap_cur;
  int snap_n_err, res_epoch;  // res_epoch: launches of the resident kernel on this workspace
  unsigned gbar_count, gbar_gen;  // grid barrier of the fused-iteration sweep (count returns to 0 after every use)
  float snap_err, pad3, pad4, pad5;
  // range of the scaled row potentials fs; slot [it & 1] is valid when `it` iterations are complete,
  // the sweep of iteration it+1 fills slot [(it+1) & 1].  lo > hi means "unknown".
  float fs_lo[2], fs_hi[2];
  // fp32-floor stop (b200ot_params::floor_patience)
  int floor_patience, stall, floor_hit, pad6;
  float best_err, b_l1, b_l2sq, pad7;
};

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) > v) {
    const int prev = atomicCAS(a, old, __float_as_int(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int prev = atomicCAS(a, old, __float_as_int(v));
    if (prev == old) break;
    old = prev;
  }
}
static_assert(sizeof(State) <= 256, "state block");

struct WsLayout {
  size_t state, err_hist, errpart, fs, gs0, gs1, a, b, log2b, snap_fs, snap_gs, part_sum, part_max,
      res_ll, res_ll_bytes, total;
  size_t m_pad, n_pad;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline WsLayout ws_layout(int n, int m) {
  WsLayout L;
  L.m_pad = align_up((size_t)m, 64);
  L.n_pad = align_up((size_t)n, 64);
  size_t off = 0;
  L.state = off;
  off += 256;
  L.err_hist = off;
  off += kErrHistCap * sizeof(float);
  L.errpart = off;
  off += 4096 * sizeof(double);
  auto vec = [&](size_t elems) {
    size_t o = off;
    off += align_up(elems * sizeof(float), 256);
    return o;
  };
  L.fs = vec(L.n_pad);
  L.gs0 = vec(L.m_pad);
  L.gs1 = vec(L.m_pad);
  L.a = vec(L.n_pad);
  L.b = vec(L.m_pad);
  L.log2b = vec(L.m_pad);
  L.snap_fs = vec(L.n_pad);
  L.snap_gs = vec(L.m_pad);
  L.part_sum = vec((size_t)kNpCap * L.m_pad);
  L.part_max = vec((size_t)kNpCap * L.m_pad);
  // tagged 64-bit words of the resident kernel: column partials [kResMaxCtas][m_pad], g [m_pad], error [kResMaxCtas][2]
  L.res_ll = off;
  L.res_ll_bytes = m <= kResMaxM ? ((size_t)kResMaxCtas * L.m_pad + L.m_pad + 2 * (size_t)kResMaxCtas) * 8 : 0;
  off += align_up(L.res_ll_bytes, 256);
  L.total = off;
  return L;
}

struct WsPtrs {
  State* st;
  float* err_hist;
  double* errpart;
  float *fs, *gs0, *gs1, *a, *b, *log2b, *snap_fs, *snap_gs, *part_sum, *part_max;
  void* res_ll;  // null when the rows are too wide for the resident kernel
  size_t m_pad;
};
static inline WsPtrs ws_ptrs(void* ws, const WsLayout& L) {
  char* p = static_cast<char*>(ws);
  WsPtrs w;
  w.st = reinterpret_cast<State*>(p + L.state);
  w.err_hist = reinterpret_cast<float*>(p + L.err_hist);
  w.errpart = reinterpret_cast<double*>(p + L.errpart);
  w.fs = reinterpret_cast<float*>(p + L.fs);
  w.gs0 = reinterpret_cast<float*>(p + L.gs0);
  w.gs1 = reinterpret_cast<float*>(p + L.gs1);
  w.a = reinterpret_cast<float*>(p + L.a);
  w.b = reinterpret_cast<float*>(p + L.b);
  w.log2b = reinterpret_cast<float*>(p + L.log2b);
  w.snap_fs = reinterpret_cast<float*>(p + L.snap_fs);
  w.snap_gs = reinterpret_cast<float*>(p + L.snap_gs);
  w.part_sum = reinterpret_cast<float*>(p + L.part_sum);
  w.part_max = reinterpret_cast<float*>(p + L.part_max);
  w.res_ll = L.res_ll_bytes ? static_cast<void*>(p + L.res_ll) : nullptr;
  w.m_pad = L.m_pad;
  return w;
}

// The solver's state machine after one completed iteration (shared by finalize_kernel and the resident
// kernel, which evolves a private copy per CTA): iteration count, error history, the reference stopping rule,
// the fp32-floor rule, and which g buffer is current.  `err_hist` may be null (only one writer records it).
__device__ __forceinline__ void advance_state(State& s, float err, bool range_untracked, float* err_hist) {
  const int norm = s.err_norm;
  const int it = s.it + 1;
  s.it = it;
  // fs range bookkeeping: the sweep that just ran filled slot [it & 1] (the two-sweep path does not track
  // it: mark unknown); open the other slot for the next sweep
  if (range_untracked) {
    s.fs_lo[it & 1] = INFINITY;
    s.fs_hi[it & 1] = -INFINITY;
  }
  s.fs_lo[(it + 1) & 1] = INFINITY;
  s.fs_hi[(it + 1) & 1] = -INFINITY;
  const int ce = s.check_every;
  const bool check = (it % ce) == (s.check_phase % ce);
  bool stop = false;
  if (check) {
    s.err = err;
    const int ne = s.n_err;
    if (err_hist && ne < kErrHistCap) err_hist[ne] = err;
    s.n_err = ne + 1;
    stop = s.stop_inclusive ? (err <= s.tol) : (err < s.tol);
    if (!stop && s.floor_patience > 0 && s.tol > 0.f) {  // tol == 0 asks for a fixed iteration count
      // resolution floor: no new minimum for `patience` checks, and already far below |b|
      const float scale = norm == B200OT_NORM_L1 ? s.b_l1 * 1e-4f
                          : norm == B200OT_NORM_L2 ? sqrtf(s.b_l2sq) * 1e-4f : s.b_l2sq * 1e-8f;
      if (err < s.best_err * 0.999f) {
        s.best_err = err;
        s.stall = 0;
      } else if (++s.stall >= s.floor_patience && err <= scale) {
        stop = true;
        s.floor_hit = 1;
      }
    }
  }
  if (stop) {
    s.converged = 1;
    s.done = 1;
  } else if (it >= s.max_iter) {
    if (!check) s.err = err;
    s.done = 1;
  } else {
    s.cur ^= 1;
  }
}

// resident.cu: queue `iters` iterations as one persistent cooperative launch.  0 = queued, 1 = not applicable
// to this problem (caller uses the launch-per-sweep kernels), < 0 = error.
int resident_try_enqueue(const float* C, int ldc, int n, int m, int iters, const WsPtrs& w, cudaStream_t s);
bool resident_describe(int n, int m, char* buf, int buf_len);
bool resident_applicable(int n, int m);

}  // namespace b200ot
