// Resident Sinkhorn solver: a whole chunk of iterations in ONE persistent cooperative launch.
//
// Same mathematics as the single-sweep kernels of sinkhorn.cu (one evaluation of
// t_ij = 2^(fs_i + gs_j - k C_ij) per element and iteration gives the f update, the column marginal of the
// new plan -- i.e. the convergence check of the reference, perturbot/perturbot/match/utils.py:80-89 -- and
// the next g update), but built for problems whose iteration is short (n*m up to 8192^2: BASELINE configs
// C1/C3, the reference-native 512^2 / 2048^2 feature problems of MRI_PET_OT_nojax.py:91-145): there the
// per-iteration cost of the launch-per-sweep path is launches, ramp-up and the finalize kernel, not bytes.
//
//  * one CTA per SM owns a contiguous block of rows (whole rows, no cluster exchange); thread t owns the
//    column quads t, t+T, ...: column accumulators and the scaled g live in registers for a whole sweep;
//  * the rows sit in a shared-memory ring filled by TMA bulk copies.  When the block fits the ring
//    (n*m*4 <= ~28 MB, e.g. 2048^2) C is read from HBM ONCE per launch and every later iteration runs out of
//    shared memory.  Otherwise the ring is a cache of the most recent groups and sweeps alternate direction
//    ("snake"): the tail of one sweep is the head of the next, so the ring content -- and whatever of the
//    block is still in L2 -- is reused instead of re-fetched;
//  * the column reduction across CTAs, the marginal error, the next g and the stopping rule run inside the
//    kernel behind two grid barriers per iteration: partials -> barrier -> CTA c folds its slice of the
//    columns over all CTAs in fixed order (bit-reproducible), writes g and one error partial -> barrier ->
//    every CTA folds the same error partials and advances a private copy of the state machine, so all CTAs
//    take identical decisions without another exchange.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "solver_state.cuh"

namespace b200ot {

constexpr int kResMaxStages = 16;
constexpr size_t kResSmemMax = 232448 - 1024;

struct ResidentArgs {
  const float* C;
  long long ldc;
  int n, m;
  State* st;
  float* fs;
  float* gs0;
  float* gs1;
  const float* a;
  const float* b;
  const float* log2b;
  float* part;  // [grid][stride] column partials
  size_t stride;
  double* errpart;  // [grid]
  float* err_hist;
  unsigned* gbar;   // zeroed by the host before the launch
  int iters;        // iteration budget of this launch (>= 1)
  int ng;           // ring depth in row groups
  int rows_cap;     // rows per CTA, rounded up to whole groups
  int snake;        // alternate the sweep direction when the block does not fit the ring
  int evict_first;  // L2 evict-first hint on the bulk copies
  int red_groups;   // thread groups of the slice fold (power of two)
};

// All CTAs are co-resident (cooperative launch).  Monotonic counter: barrier k completes at k * grid.
// bar.sync + red.release.gpu / ld.acquire.gpu + bar.sync (the CUTLASS semaphore pattern): the release is
// cumulative over the CTA's writes ordered before it by the block barrier, so no extra fences (a
// __threadfence() on each side, MEMBAR.SC.GPU, cost ~4 us per iteration when measured).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000ll) __trap();  // a lost CTA traps instead of hanging the GPU box
    }
  }
  __syncthreads();
}

// Sum R per-lane values over the warp with (R - 1) + 5 - log2(R) shuffles instead of 5 R: each of the first
// log2(R) butterfly steps halves the number of values a lane carries.  Returns, in every lane, the total of row
// lane >> (5 - log2 R).  Fixed order: deterministic.
template <int R>
__device__ __forceinline__ float warp_sum_rows(float (&v)[R], int lane) {
  static_assert(R == 1 || R == 2 || R == 4 || R == 8, "rows per group");
  int off = 16;
#pragma unroll
  for (int h = R / 2; h >= 1; h >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = upper ? v[i] : v[i + h];
      const float keep = upper ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  float x = v[0];
  for (; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

template <int T, int NCH, int R, bool PIPE>
__global__ void __launch_bounds__(T, 1) resident_kernel(const ResidentArgs p) {
  constexpr int CPT = 4 * NCH;
  constexpr int W = T * CPT;  // floats per staged row
  constexpr int NW = T / 32;
  constexpr int LPR = 32 / R;        // lanes per row in the warp-level reductions
  constexpr int V = R * NW / 32;     // warp partials each lane folds in the second stage
  static_assert(R * NW >= 32 && (R * NW) % 32 == 0 && V <= 4, "second-stage layout");
  extern __shared__ __align__(128) unsigned char smem[];

  if (p.st->done) return;  // grid-uniform

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int NG = p.ng;

  float* stage = reinterpret_cast<float*>(smem);  // [NG][R][W]
  unsigned char* q = smem + (size_t)NG * R * W * sizeof(float);
  uint64_t* full = reinterpret_cast<uint64_t*>(q);
  q += kResMaxStages * 8;
  State* ls = reinterpret_cast<State*>(q);
  q += 256;
  double* dred = reinterpret_cast<double*>(q);
  q += 32 * 8;
  int* fillcnt = reinterpret_cast<int*>(q);
  q += kResMaxStages * 4;
  int* bad_sm = reinterpret_cast<int*>(q);
  q += 16;
  float* red = reinterpret_cast<float*>(q);  // [3][R][NW] warp partials of the row sums (3 groups in flight)
  q += 3 * R * NW * 4;
  float* shs = reinterpret_cast<float*>(q);  // [T] slice-fold scratch
  q += T * 4;
  float* fs_sm = reinterpret_cast<float*>(q);  // [rows_cap] scaled row potentials of this CTA's rows
  q += (size_t)p.rows_cap * 4;
  float* a_sm = reinterpret_cast<float*>(q);  // [rows_cap]

  // rows of this CTA: whole groups of R rows, contiguous
  const int ngroups = (p.n + R - 1) / R;
  const int g_begin = (int)((long long)cta * ngroups / G);
  const int g_end = (int)((long long)(cta + 1) * ngroups / G);
  const int cnt = g_end - g_begin;
  const int row_base = g_begin * R;
  const bool streaming = cnt > NG;

  if (tid < (int)(sizeof(State) / 4)) reinterpret_cast<int*>(ls)[tid] = reinterpret_cast<const int*>(p.st)[tid];
  for (int i = tid; i < cnt * R; i += T) {
    const int row = row_base + i;
    const bool ok = row < p.n;
    fs_sm[i] = ok ? p.fs[row] : -INFINITY;  // rows past n: t = 2^-inf = 0, weight 0
    a_sm[i] = ok ? p.a[row] : 0.f;
  }
  const uint64_t pol = p.evict_first ? policy_evict_first() : 0ull;
  const uint32_t row_bytes = (uint32_t)p.m * 4u;
  auto issue = [&](int gi) {  // thread 0 only
    const int s = gi % NG;
    const uint32_t bar = smem_u32(full + s);
    fillcnt[s] += 1;
    mbar_arrive_expect_tx(bar, row_bytes * R);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int row = row_base + gi * R + r;
      row = row < p.n ? row : p.n - 1;  // ragged last group: finite data, weight 0
      const float* src = p.C + (long long)row * p.ldc;
      const uint32_t dst = smem_u32(stage + ((size_t)s * R + r) * W);
      if (p.evict_first)
        bulk_g2s_hint(dst, src, row_bytes, bar, pol);
      else
        bulk_g2s(dst, src, row_bytes, bar);
    }
  };
  if (tid == 0) {
    for (int s = 0; s < NG; ++s) {
      mbar_init(smem_u32(full + s), 1);
      fillcnt[s] = 0;
    }
    *bad_sm = 0;
    fence_mbar_init();
    const int pre = cnt < NG ? cnt : NG;
    for (int i = 0; i < pre; ++i) issue(i);
  }
  __syncthreads();

  const float k = ls->kscale;
  const int norm = ls->err_norm;
  unsigned nbar = 0;

  // column slice this CTA folds after the sweep
  const int CB = (p.m + G - 1) / G;
  const int j0 = cta * CB;
  int ncol = p.m - j0;
  ncol = ncol < 0 ? 0 : (ncol > CB ? CB : ncol);
  const int groups = p.red_groups, CBT = T / groups;
  const int rgrp = tid / CBT, rc = tid - rgrp * CBT;

  const bool last_ok = ((NCH - 1) * T + tid) * 4 < p.m;
  const bool last_any = __any_sync(0xffffffffu, last_ok);

  float gsv[CPT], acc[CPT];
  auto load_g = [&](const float* gsrc) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (c * T + tid) * 4;
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < p.m) g4 = __ldcg(reinterpret_cast<const float4*>(gsrc + col));  // written by peers: L2, not L1
      gsv[c * 4 + 0] = g4.x;
      gsv[c * 4 + 1] = g4.y;
      gsv[c * 4 + 2] = g4.z;
      gsv[c * 4 + 3] = g4.w;
    }
  };
  load_g(ls->cur ? p.gs1 : p.gs0);
  int bad = 0;  // sticky: reported with the next error fold

  for (int li = 0; li < p.iters; ++li) {
    if (ls->done) break;  // identical on every CTA
    const int cur = ls->cur;
    const float* gs = cur ? p.gs1 : p.gs0;
    float* gnext = cur ? p.gs0 : p.gs1;
    // the marginal error is only consumed on check iterations, at max_iter, and (to report a lost sum before
    // the host looks) on the last iteration of the launch
    const int itn = ls->it + 1;
    const bool need_err = (itn % ls->check_every) == (ls->check_phase % ls->check_every) ||
                          itn >= ls->max_iter || li + 1 == p.iters;
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = 0.f;

    const bool fwd = !(p.snake && streaming && (li & 1));
    auto gidx = [&](int pos) { return fwd ? pos : cnt - 1 - pos; };

    // P1 of the group at sweep position pos: exponentials into registers, row sums warp -> red[pos % 3]
    auto front = [&](int pos, float (&t)[R][CPT]) {
      const int gi = gidx(pos);
      const int s = gi % NG;
      const int lr0 = gi * R;
      mbar_wait(smem_u32(full + s), (uint32_t)((fillcnt[s] - 1) & 1));
      float ps[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* srow = stage + ((size_t)s * R + r) * W + tid * 4;
        const float sh = fs_sm[lr0 + r];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < NCH - 1 || last_any) {
            const float4 v = *reinterpret_cast<const float4*>(srow + c * (T * 4));
            t[r][c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0] + sh));
            t[r][c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1] + sh));
            t[r][c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2] + sh));
            t[r][c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3] + sh));
          }
        }
        if (!last_ok) t[r][CPT - 4] = t[r][CPT - 3] = t[r][CPT - 2] = t[r][CPT - 1] = 0.f;  // past m: stale smem
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          sum += (t[r][c * 4 + 0] + t[r][c * 4 + 1]) + (t[r][c * 4 + 2] + t[r][c * 4 + 3]);
        ps[r] = sum;
      }
      const float v = warp_sum_rows<R>(ps, lane);
      if ((lane & (LPR - 1)) == 0) red[((pos % 3) * R + lane / LPR) * NW + warp] = v;
    };

    // P2: fold the warp partials, w_i = a_i / r_i, new row potential, column accumulators
    auto back = [&](int pos, float (&t)[R][CPT]) {
      const int lr0 = gidx(pos) * R;
      // lane l folds V consecutive warp partials of row l / LPR, then LPR lanes are combined
      float x = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) x += red[(pos % 3) * R * NW + lane * V + i];
#pragma unroll
      for (int off = LPR / 2; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
      float wr[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float rt = __shfl_sync(0xffffffffu, x, r * LPR);
        const float ar = a_sm[lr0 + r];
        const bool live = ar > 0.f;
        wr[r] = live ? __fdividef(ar, rt) : 0.f;
        if (tid == r && row_base + lr0 + r < p.n) {
          const float fnew = live ? fs_sm[lr0 + r] + (log2f(ar) - log2f(rt)) : -INFINITY;
          fs_sm[lr0 + r] = fnew;
          p.fs[row_base + lr0 + r] = fnew;
          if (live && !(fabsf(fnew) < INFINITY)) bad = 1;  // vanished / overflowed row sum
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = fmaf(t[r][c], wr[r], acc[c]);
    };

    // after the block barrier that follows front(pos + 1): the stages of positions <= pos + 1 are drained
    auto refill = [&](int pos) {
      if (tid != 0 || !streaming) return;
      auto one = [&](int d) {
        if (d >= cnt) return;
        const int gi = gidx(d);
        const int nxt = fwd ? gi + NG : gi - NG;
        if (nxt >= 0 && nxt < cnt) {
          fence_proxy_async();
          issue(nxt);
        }
      };
      if (PIPE) {
        if (pos == 0) one(0);
        one(pos + 1);
      } else {
        one(pos);
      }
    };

    if (PIPE) {
      // software pipeline over groups: the reduction chain of group pos (shuffles, block barrier, divide)
      // overlaps the exponentials of group pos + 1.  Two register sets; red[] is triple-buffered because
      // front(pos + 2) may run while a slower warp is still in back(pos).
      float tA[R][CPT], tB[R][CPT];
      front(0, tA);
      for (int pos = 0; pos < cnt; pos += 2) {
        if (pos + 1 < cnt) front(pos + 1, tB);
        __syncthreads();
        refill(pos);
        back(pos, tA);
        if (pos + 1 < cnt) {
          if (pos + 2 < cnt) front(pos + 2, tA);
          __syncthreads();
          refill(pos + 1);
          back(pos + 1, tB);
        }
      }
    } else {
      float tA[R][CPT];
      for (int pos = 0; pos < cnt; ++pos) {
        front(pos, tA);
        __syncthreads();
        refill(pos);
        back(pos, tA);
      }
    }

    // ---- column partials of this CTA -> global ---------------------------------------------------------
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (c * T + tid) * 4;
      if (col < p.m)
        *reinterpret_cast<float4*>(p.part + (size_t)cta * p.stride + col) =
            make_float4(acc[c * 4 + 0], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
    }
    grid_barrier(p.gbar, (unsigned)G * (++nbar));

    // ---- fold this CTA's column slice over all CTAs (fixed order), next g, error partial ----------------
    double e = 0.0;
    for (int jj0 = 0; jj0 < CB; jj0 += CBT) {  // uniform trip count
      const int jj = jj0 + rc;
      const int j = j0 + jj;
      float pa = 0.f;
      if (jj < ncol)
        for (int pp = rgrp; pp < G; pp += groups) pa += __ldcg(p.part + (size_t)pp * p.stride + j);
      shs[tid] = pa;
      __syncthreads();
      if (rgrp == 0 && jj < ncol) {
        float sj = 0.f;
        for (int g2 = 0; g2 < groups; ++g2) sj += shs[g2 * CBT + rc];
        const float l2s = log2f(sj);
        const float bj = p.b[j];
        if (need_err) {
          const double d = (double)sj - (double)bj;
          e += (norm == B200OT_NORM_L1) ? fabs(d) : d * d;
        }
        const float gn = bj > 0.f ? __ldcg(gs + j) + (p.log2b[j] - l2s) : -INFINITY;  // b_j = 0: v_j = 0
        gnext[j] = gn;
        if (bj > 0.f && !(fabsf(gn) < INFINITY)) bad = 1;
      }
      if (jj0 + CBT < CB) __syncthreads();  // shs is reused by the next pass
    }
    if (need_err) {
      e = warp_sum(e);
      if (lane == 0) dred[warp] = e;
      const int anybad = __syncthreads_or(bad);
      if (tid == 0) {
        double tot = 0.0;
        for (int w2 = 0; w2 < NW; ++w2) tot += dred[w2];
        // a lost sum anywhere poisons the error every CTA folds next: all of them stop together
        p.errpart[cta] = anybad ? (double)NAN : tot;
      }
    }
    grid_barrier(p.gbar, (unsigned)G * (++nbar));

    // ---- next g into registers; every CTA folds the same error partials and advances its state copy ----
    load_g(gnext);
    if (need_err) {
      double ep = 0.0;
      for (int i = tid; i < G; i += T) ep += __ldcg(p.errpart + i);
      ep = warp_sum(ep);
      if (lane == 0) dred[warp] = ep;
      __syncthreads();
    }
    if (tid == 0) {
      double tot = 0.0;
      if (need_err)
        for (int w2 = 0; w2 < NW; ++w2) tot += dred[w2];
      if (!(tot == tot)) {  // fast path lost a sum: stop here, the host rewinds and replays robustly
        ls->bad = 1;
        ls->done = 1;
      } else {
        const float err = (norm == B200OT_NORM_L2) ? (float)sqrt(tot) : (float)tot;
        advance_state(*ls, err, true, cta == 0 ? p.err_hist : nullptr);
      }
      // one-directional streaming: the ring holds the tail of this sweep, refill it with the head of the next
      if (streaming && !p.snake && !ls->done && li + 1 < p.iters) {
        fence_proxy_async();
        for (int i = 0; i < NG; ++i) issue(i);
      }
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) *p.st = *ls;
}

// ---- host side ----------------------------------------------------------------------------------------
struct ResCfg {
  int T, NCH, R, G, NG, rows_cap, red_groups, cnt_max;
  size_t smem;
};

static bool resident_pick(int n, int m, ResCfg* c) {
  if (n < 1 || m < 4 || (m & 3) || m > 8192) return false;
  const int T = m <= 512 ? 128 : m <= 1024 ? 256 : 512;
  const int NCH = (m + 4 * T - 1) / (4 * T);
  const int R = NCH == 1 ? 8 : NCH == 2 ? 4 : 2;
  const int ngroups = (n + R - 1) / R;
  int G = sm_count();
  if (G < 1) return false;
  if (G > kNpCap) G = kNpCap;
  if (G > ngroups) G = ngroups;
  const int cnt_max = (ngroups + G - 1) / G;
  const int rows_cap = cnt_max * R;
  const size_t stage = (size_t)R * T * 4 * NCH * sizeof(float);
  const size_t fixed = kResMaxStages * 8 + 256 + 32 * 8 + kResMaxStages * 4 + 16 + 3 * R * (T / 32) * 4 +
                       (size_t)T * 4 + 2 * (size_t)rows_cap * 4 + 128;
  if (fixed + 2 * stage > kResSmemMax) return false;
  int NG = (int)((kResSmemMax - fixed) / stage);
  if (NG > kResMaxStages) NG = kResMaxStages;
  if (NG > cnt_max) NG = cnt_max;
  if (NG < 2 && cnt_max > NG) return false;
  const int CB = (m + G - 1) / G;
  int groups = 1;
  while (groups * 2 <= 16 && T / (groups * 2) >= CB) groups *= 2;
  c->T = T;
  c->NCH = NCH;
  c->R = R;
  c->G = G;
  c->NG = NG;
  c->rows_cap = rows_cap;
  c->red_groups = groups;
  c->cnt_max = cnt_max;
  c->smem = (size_t)NG * stage + fixed;
  return true;
}

template <int T, int NCH, int R, bool PIPE>
static cudaError_t resident_launch(const ResidentArgs& a, int G, size_t smem, cudaStream_t s) {
  static bool attr_set = false;  // one flag per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(resident_kernel<T, NCH, R, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kResSmemMax);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)G);
  cfg.blockDim = dim3((unsigned)T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch fails
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, resident_kernel<T, NCH, R, PIPE>, a);
}

// auto rule: iterations short enough that launches and the finalize kernel dominate the sweep
static bool resident_wanted(int n, int m) {
  const char* e = getenv("B200OT_RESIDENT");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return (double)n * (double)m <= 8192.0 * 8192.0;
}

static bool g_resident_broken = false;  // a failed cooperative launch disables the path for the process

int resident_try_enqueue(const float* C, int ldc, int n, int m, int iters, const WsPtrs& w, cudaStream_t s) {
  if (iters < 1 || g_resident_broken || !resident_wanted(n, m)) return 1;
  if ((ldc & 3) || (reinterpret_cast<uintptr_t>(C) & 15)) return 1;
  ResCfg c;
  if (!resident_pick(n, m, &c)) return 1;
  ResidentArgs a;
  a.C = C;
  a.ldc = ldc;
  a.n = n;
  a.m = m;
  a.st = w.st;
  a.fs = w.fs;
  a.gs0 = w.gs0;
  a.gs1 = w.gs1;
  a.a = w.a;
  a.b = w.b;
  a.log2b = w.log2b;
  a.part = w.part_sum;
  a.stride = w.m_pad;
  a.errpart = w.errpart;
  a.err_hist = w.err_hist;
  a.gbar = w.gbar;
  a.iters = iters;
  a.ng = c.NG;
  a.rows_cap = c.rows_cap;
  const char* es = getenv("B200OT_RES_SNAKE");
  a.snake = (es && es[0] == '0') ? 0 : 1;
  const char* ee = getenv("B200OT_RES_EVICT");
  a.evict_first = (ee && ee[0] == '1') ? 1 : 0;
  a.red_groups = c.red_groups;
  B200OT_CUDA_OK(cudaMemsetAsync(w.gbar, 0, sizeof(unsigned), s));
  cudaError_t e = cudaSuccess;
  const char* ep = getenv("B200OT_RES_PIPE");
  const bool pipe = !(ep && ep[0] == '0');
#define B200OT_RES(T_, N_, R_) \
  e = pipe ? resident_launch<T_, N_, R_, true>(a, c.G, c.smem, s) : resident_launch<T_, N_, R_, false>(a, c.G, c.smem, s)
  if (c.T == 128)
    B200OT_RES(128, 1, 8);
  else if (c.T == 256)
    B200OT_RES(256, 1, 8);
  else if (c.NCH == 1)
    B200OT_RES(512, 1, 8);
  else if (c.NCH == 2)
    B200OT_RES(512, 2, 4);
  else if (c.NCH == 3)
    B200OT_RES(512, 3, 2);
  else
    B200OT_RES(512, 4, 2);
#undef B200OT_RES
  if (e != cudaSuccess) {  // e.g. cooperative launch not possible here: use the launch-per-sweep path
    set_last_cuda_error(e, "resident_kernel launch (falling back to per-sweep launches)");
    (void)cudaGetLastError();
    g_resident_broken = true;
    return 1;
  }
  return 0;
}

bool resident_applicable(int n, int m) {
  ResCfg c;
  return !g_resident_broken && resident_wanted(n, m) && resident_pick(n, m, &c);
}

bool resident_describe(int n, int m, char* buf, int buf_len) {
  ResCfg c;
  if (g_resident_broken || !resident_wanted(n, m) || !resident_pick(n, m, &c)) return false;
  const char* es = getenv("B200OT_RES_SNAKE");
  const bool snake = !(es && es[0] == '0');
  snprintf(buf, buf_len,
           "resident_kernel (persistent, cooperative): %d CTAs x %d threads, %d cols/thread, %d rows/group, "
           "<=%d groups/CTA, ring %d x %zu B (%s), 2 grid barriers/iteration, smem %zu B/CTA",
           c.G, c.T, 4 * c.NCH, c.R, c.cnt_max, c.NG, (size_t)c.R * c.T * 4 * c.NCH * 4,
           c.cnt_max <= c.NG ? "C shared-memory resident" : (snake ? "snake sweeps" : "forward sweeps"), c.smem);
  return true;
}

}  // namespace b200ot
