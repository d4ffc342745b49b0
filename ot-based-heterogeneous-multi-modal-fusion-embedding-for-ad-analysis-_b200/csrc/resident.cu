// Resident Sinkhorn solver: a whole chunk of iterations in ONE persistent cooperative launch.
//
// Same mathematics as the single-sweep kernels of sinkhorn.cu (one evaluation of
// t_ij = 2^(fs_i + gs_j - k C_ij) per element and iteration gives the f update, the column marginal of the
// new plan -- i.e. the convergence check of the reference, perturbot/perturbot/match/utils.py:80-89 -- and
// the next g update), but built for problems whose iteration is short (n*m up to 8192^2: BASELINE configs
// C1/C3, the reference-native 512^2 / 2048^2 feature problems of MRI_PET_OT_nojax.py:91-145): there the
// per-iteration cost of the launch-per-sweep path is launches, ramp-up and the finalize kernel, not bytes.
//
//  * two 256-thread CTAs per SM (one 512-thread CTA for rows wider than 4096 columns) each own a contiguous block
//    of whole rows (no cluster exchange); thread t owns the column quads t, t+T, ...: column accumulators and
//    the scaled g live in registers for a whole sweep;
//  * the rows sit in a shared-memory ring filled by TMA bulk copies.  When the block fits the ring
//    (n*m*4 <= ~28 MB, e.g. 2048^2) C is read from HBM ONCE per launch and every later iteration runs out of
//    shared memory.  Otherwise the ring is a cache of the most recent groups and sweeps alternate direction
//    ("snake"): the tail of one sweep is the head of the next, so the ring content -- and whatever of the
//    block is still in L2 -- is reused instead of re-fetched;
//  * the column reduction across CTAs, the marginal error, the next g and the stopping rule run inside the
//    kernel WITHOUT a grid barrier: every value that crosses CTAs is a tagged 64-bit word that its consumers
//    poll (see "tagged words" below).  CTA c folds its slice of the columns over all CTAs in fixed order
//    (bit-reproducible) and publishes g and, on check iterations, one error partial; every CTA gathers g, folds
//    the same error partials and advances a private copy of the state machine, so all CTAs take identical
//    decisions without another exchange.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "solver_state.cuh"

namespace b200ot {

constexpr int kResMaxStages = 16;
constexpr size_t kResSmemMax = 113 * 1024;       // two CTAs per SM (T <= 256)
constexpr size_t kResSmemMaxWide = 232448 - 1024;  // one 512-thread CTA per SM (rows wider than 4096 columns)

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
  unsigned long long* part64;  // [grid][stride] column partials, each word = {value, sequence tag}
  size_t stride;
  unsigned long long* g64;    // [m] next scaled g, tagged
  unsigned long long* err64;  // [grid][2] error partial (the two halves of a double), tagged
  float* err_hist;
  int iters;        // iteration budget of this launch (1 .. kResMaxItersPerLaunch)
  int ng;           // ring depth in row groups
  int rows_cap;     // rows per CTA, rounded up to whole groups
  int snake;        // alternate the sweep direction when the block does not fit the ring
  int evict_first;  // L2 evict-first hint on the bulk copies
  int red_groups;   // thread groups of the slice fold (power of two)
  int spin_ns;      // back-off between polls of a word that is not there yet
};

// ---- tagged words: the exchange between CTAs needs no barrier and no fence ------------------------------------
// Every value that crosses CTAs travels in ONE naturally aligned 64-bit word {fp32 value, 32-bit sequence tag}
// (the NCCL "LL" idea): a 64-bit access is single-copy atomic, so a consumer that sees the tag of iteration k
// has the value of iteration k.  Consumers poll the word itself.  Tags are (launch epoch, iteration in launch);
// the epoch lives in the state block and b200ot_sinkhorn_setup zeroes the buffers, so a stale word never matches.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack_tag(float v, unsigned tag) { return ((u64)tag << 32) | (u64)__float_as_uint(v); }
__device__ __forceinline__ u64 pack_tag_u(unsigned v, unsigned tag) { return ((u64)tag << 32) | (u64)v; }
__device__ __forceinline__ u64 ld_poll(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ld_poll2(const u64* p, u64& a, u64& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_tag(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_tag2(u64* p, u64 a, u64 b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ bool tag_is(u64 v, unsigned tag) { return (unsigned)(v >> 32) == tag; }
// a producer that never shows up traps instead of hanging the GPU box
__device__ __forceinline__ void spin_guard(long long& t0, int ns) {
  if (t0 == 0)
    t0 = clock64();
  else if (clock64() - t0 > 4000000000ll)
    __trap();
  __nanosleep(ns);
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
#pragma unroll
  for (int o = 16 / R; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int T, int NCH, int R>
__global__ void __launch_bounds__(T, (T > 256 ? 1 : 2)) resident_kernel(const ResidentArgs p) {
  constexpr int CPT = 4 * NCH;
  constexpr int W = T * CPT;  // floats per staged row
  constexpr int NW = T / 32;
  constexpr int LPR = 32 / R;                  // lanes per row in the warp-level reductions
  constexpr int V = (NW * R + 31) / 32;        // warp partials each lane folds in the second stage
  constexpr int kFold = NCH <= 4 ? 10 : 8;      // tagged words a thread polls at once in the slice fold
  constexpr int kGq = NCH <= 4 ? NCH : 2;      // column quads a thread polls at once when it gathers g
  static_assert(NW <= 32 && (NW & (NW - 1)) == 0 && R * CPT <= 32, "register budget: R * CPT exponentials per thread and group");
  extern __shared__ __align__(128) unsigned char smem[];

  if (p.st->done) return;  // grid-uniform

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int NG = p.ng;

  // fixed-size arrays first (compile-time offsets), then the per-row vectors, then the ring
  unsigned char* q = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(q);
  q += kResMaxStages * 8;
  State* ls = reinterpret_cast<State*>(q);
  q += 256;
  double* dred = reinterpret_cast<double*>(q);
  q += 32 * 8;
  int* fillcnt = reinterpret_cast<int*>(q);
  q += kResMaxStages * 4;
  float* red = reinterpret_cast<float*>(q);  // [2][R][NW] warp partials of the row sums
  q += 2 * R * NW * 4;
  float* shs = reinterpret_cast<float*>(q);  // [T] slice-fold scratch
  q += T * 4;
  float* fs_sm = reinterpret_cast<float*>(q);  // [rows_cap] scaled row potentials of this CTA's rows
  q += (size_t)p.rows_cap * 4;
  float* a_sm = reinterpret_cast<float*>(q);  // [rows_cap]
  q += (size_t)p.rows_cap * 4;
  float* la_sm = reinterpret_cast<float*>(q);  // [rows_cap] log2 a_i
  q += (size_t)p.rows_cap * 4;
  float* stage = reinterpret_cast<float*>(smem + (((size_t)(q - smem) + 127) & ~(size_t)127));  // [NG][R][W]
  const uint32_t full_u32 = smem_u32(full);

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
    const float ai = ok ? p.a[row] : 0.f;
    a_sm[i] = ai;
    la_sm[i] = ai > 0.f ? log2f(ai) : -INFINITY;
  }
  const uint64_t pol = p.evict_first ? policy_evict_first() : 0ull;
  const uint32_t row_bytes = (uint32_t)p.m * 4u;
  auto issue = [&](int gi, int s) {  // thread 0 only: group gi into ring slot s
    const uint32_t bar = full_u32 + 8u * (uint32_t)s;
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
      mbar_init(full_u32 + 8u * (uint32_t)s, 1);
      fillcnt[s] = 0;
    }
    fence_mbar_init();
    const int pre = cnt < NG ? cnt : NG;
    for (int i = 0; i < pre; ++i) issue(i, i);  // slot of group gi is gi % NG
  }
  __syncthreads();

  const float k = ls->kscale;
  const int norm = ls->err_norm;
  const unsigned epoch = ((unsigned)ls->res_epoch + 1u) & 0xffffu;

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
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (c * T + tid) * 4;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < p.m) g4 = *reinterpret_cast<const float4*>((ls->cur ? p.gs1 : p.gs0) + col);
    gsv[c * 4 + 0] = g4.x;
    gsv[c * 4 + 1] = g4.y;
    gsv[c * 4 + 2] = g4.z;
    gsv[c * 4 + 3] = g4.w;
  }
  int bad = 0;  // sticky: reported with the next error fold
  const int s_last = (cnt - 1) % NG;

  for (int li = 0; li < p.iters; ++li) {
    if (ls->done) break;  // identical on every CTA
    const int cur = ls->cur;
    float* gnext = cur ? p.gs0 : p.gs1;
    const unsigned tag = (epoch << 16) | (unsigned)(li + 1);
    // the marginal error is only consumed on check iterations, at max_iter, and (to report a lost sum before
    // the host looks) on the last iteration of the launch
    const int itn = ls->it + 1;
    const bool need_err = (itn % ls->check_every) == (ls->check_phase % ls->check_every) ||
                          itn >= ls->max_iter || li + 1 == p.iters;
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = 0.f;

    // ---- sweep over this CTA's row groups ---------------------------------------------------------------
    const bool fwd = !(p.snake && streaming && (li & 1));
    int gi = fwd ? 0 : cnt - 1;
    int s = fwd ? 0 : s_last;
    for (int pos = 0; pos < cnt; ++pos) {
      const int lr0 = gi * R;
      mbar_wait(full_u32 + 8u * (uint32_t)s, (uint32_t)((fillcnt[s] - 1) & 1));
      float t[R][CPT], ps[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* srow = stage + ((size_t)s * R + r) * W + tid * 4;
        const float sh = fs_sm[lr0 + r];
#pragma unroll
        for (int c = 0; c < NCH - 1; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(srow + c * (T * 4));
          t[r][c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0] + sh));
          t[r][c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1] + sh));
          t[r][c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2] + sh));
          t[r][c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3] + sh));
        }
      }
      if (last_any) {  // the last quad may lie past m: stale shared memory there, masked after the fact
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float* srow = stage + ((size_t)s * R + r) * W + tid * 4;
          const float sh = fs_sm[lr0 + r];
          constexpr int c = NCH - 1;
          const float4 v = *reinterpret_cast<const float4*>(srow + c * (T * 4));
          t[r][c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0] + sh));
          t[r][c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1] + sh));
          t[r][c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2] + sh));
          t[r][c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3] + sh));
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!last_ok) t[r][CPT - 4] = t[r][CPT - 3] = t[r][CPT - 2] = t[r][CPT - 1] = 0.f;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          sum += (t[r][c * 4 + 0] + t[r][c * 4 + 1]) + (t[r][c * 4 + 2] + t[r][c * 4 + 3]);
        ps[r] = sum;
      }
      const int par = pos & 1;
      {
        const float v = warp_sum_rows<R>(ps, lane);
        if ((lane & (LPR - 1)) == 0) red[(par * R + lane / LPR) * NW + warp] = v;
      }
      __syncthreads();  // the stage is drained by every warp; red[par] is complete
      // the per-group serial chores (ring refill, new row potentials) rotate over the warps so that no warp
      // is systematically late at the next block barrier
      if (streaming && lane == 0 && warp == ((pos + 1) & (NW - 1))) {
        const int nxt = fwd ? gi + NG : gi - NG;  // lands in the slot just drained
        if (nxt >= 0 && nxt < cnt) {
          fence_proxy_async();
          issue(nxt, s);
        }
      }
      // lane l folds V consecutive warp partials of row l / LPR, then the LPR lanes of a row are combined
      float x = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int w2 = (lane & (LPR - 1)) * V + i;
        if (w2 < NW) x += red[(par * R + lane / LPR) * NW + w2];
      }
#pragma unroll
      for (int o = LPR / 2; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      // every lane now holds the total of row lane / LPR: one divide per lane, the weights are broadcast
      const int myrow = lane / LPR;
      const float ar = a_sm[lr0 + myrow];
      const bool live = ar > 0.f;
      const float wl = live ? __fdividef(ar, x) : 0.f;
      if (warp == (pos & (NW - 1)) && (lane & (LPR - 1)) == 0 && row_base + lr0 + myrow < p.n) {
        const float fnew = live ? fs_sm[lr0 + myrow] + (la_sm[lr0 + myrow] - log2f(x)) : -INFINITY;
        fs_sm[lr0 + myrow] = fnew;
        p.fs[row_base + lr0 + myrow] = fnew;
        if (live && !(fabsf(fnew) < INFINITY)) bad = 1;  // vanished / overflowed row sum
      }
      float wr[R];
#pragma unroll
      for (int r = 0; r < R; ++r) wr[r] = __shfl_sync(0xffffffffu, wl, r * LPR);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = fmaf(t[r][c], wr[r], acc[c]);
      if (fwd) {
        ++gi;
        s = (s + 1 == NG) ? 0 : s + 1;
      } else {
        --gi;
        s = (s == 0) ? NG - 1 : s - 1;
      }
    }

    // ---- publish the column partials of this CTA (tagged words) -------------------------------------------
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (c * T + tid) * 4;
      if (col < p.m) {
        u64* dst = p.part64 + (size_t)cta * p.stride + col;
        st_tag2(dst, pack_tag(acc[c * 4 + 0], tag), pack_tag(acc[c * 4 + 1], tag));
        st_tag2(dst + 2, pack_tag(acc[c * 4 + 2], tag), pack_tag(acc[c * 4 + 3], tag));
      }
    }

    // ---- fold this CTA's column slice over all CTAs (fixed order), next g, error partial ----------------
    double e = 0.0;
    for (int jj0 = 0; jj0 < CB; jj0 += CBT) {  // uniform trip count
      const int jj = jj0 + rc;
      const int j = j0 + jj;
      float pa = 0.f;
      if (jj < ncol) {
        for (int base = rgrp; base < G; base += groups * kFold) {  // all of a thread's loads in flight at once
          u64 v[kFold];
          long long t0 = 0;
          for (;;) {
            bool ok = true;
#pragma unroll
            for (int u = 0; u < kFold; ++u) {
              const int pp = base + u * groups;
              if (pp < G) {
                v[u] = ld_poll(p.part64 + (size_t)pp * p.stride + j);
                ok = ok && tag_is(v[u], tag);
              }
            }
            if (ok) break;
            spin_guard(t0, p.spin_ns);
          }
#pragma unroll
          for (int u = 0; u < kFold; ++u)
            if (base + u * groups < G) pa += __uint_as_float((unsigned)v[u]);
        }
      }
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
        // g of the sweep that just ran: own slice, published by this thread one iteration ago (or by init)
        const float gold = (li == 0) ? (cur ? p.gs1 : p.gs0)[j] : __uint_as_float((unsigned)ld_poll(p.g64 + j));
        const float gn = bj > 0.f ? gold + (p.log2b[j] - l2s) : -INFINITY;  // b_j = 0: v_j = 0
        gnext[j] = gn;  // plain copy for the kernels that run after this launch
        st_tag(p.g64 + j, pack_tag(gn, tag));
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
        if (anybad) tot = (double)NAN;
        const u64 bits = (u64)__double_as_longlong(tot);
        st_tag2(p.err64 + 2 * (size_t)cta, pack_tag_u((unsigned)bits, tag), pack_tag_u((unsigned)(bits >> 32), tag));
      }
    }

    // ---- gather the next g (all columns) into registers ------------------------------------------------
#pragma unroll
    for (int c0 = 0; c0 < NCH; c0 += kGq) {  // kGq quads (4 words each) in flight per thread
      u64 w[kGq][4];
      long long t0 = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int cc = 0; cc < kGq; ++cc) {
          const int c = c0 + cc;
          const int col = (c * T + tid) * 4;
          if (c < NCH && col < p.m) {
            ld_poll2(p.g64 + col, w[cc][0], w[cc][1]);
            ld_poll2(p.g64 + col + 2, w[cc][2], w[cc][3]);
            ok = ok && tag_is(w[cc][0], tag) && tag_is(w[cc][1], tag) && tag_is(w[cc][2], tag) && tag_is(w[cc][3], tag);
          }
        }
        if (ok) break;
        spin_guard(t0, p.spin_ns);
      }
#pragma unroll
      for (int cc = 0; cc < kGq; ++cc) {
        const int c = c0 + cc;
        const int col = (c * T + tid) * 4;
        if (c < NCH) {
#pragma unroll
          for (int i = 0; i < 4; ++i) gsv[c * 4 + i] = col < p.m ? __uint_as_float((unsigned)w[cc][i]) : 0.f;
        }
      }
    }

    // ---- every CTA folds the same error partials and advances its copy of the state ----------------------
    if (need_err) {
      double ep = 0.0;
      for (int i = tid; i < G; i += T) {
        u64 lo, hi;
        long long t0 = 0;
        for (;;) {
          ld_poll2(p.err64 + 2 * (size_t)i, lo, hi);
          if (tag_is(lo, tag) && tag_is(hi, tag)) break;
          spin_guard(t0, p.spin_ns);
        }
        ep += __longlong_as_double((long long)(((u64)(unsigned)hi << 32) | (u64)(unsigned)lo));
      }
      ep = warp_sum(ep);
      __syncthreads();  // dred: the publisher above has read it
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
        for (int i = 0; i < NG; ++i) issue(i, i);
      }
    }
    __syncthreads();
  }
  if (cta == 0 && tid == 0) {
    ls->res_epoch = (int)epoch;
    *p.st = *ls;
  }
}

// ---- host side ----------------------------------------------------------------------------------------
constexpr int kResMaxItersPerLaunch = 60000;  // the iteration-in-launch half of the tag is 16 bits

struct ResCfg {
  int T, NCH, R, G, NG, rows_cap, red_groups, cnt_max;
  size_t smem;
};

template <int T, int NCH, int R>
static cudaError_t resident_launch(const ResidentArgs* a, int G, size_t smem, cudaStream_t s, int* occ_out) {
  static PerDeviceOnce attr_once;  // one per instantiation and device
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(resident_kernel<T, NCH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(T > 256 ? kResSmemMaxWide : kResSmemMax));
    if (e != cudaSuccess) {
      attr_once.undo();
      return e;
    }
  }
  if (!a) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_out, resident_kernel<T, NCH, R>, T, smem);
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
  return cudaLaunchKernelEx(&cfg, resident_kernel<T, NCH, R>, *a);
}

static cudaError_t resident_dispatch(int T, int NCH, const ResidentArgs* a, int G, size_t smem, cudaStream_t s,
                                     int* occ_out) {
  if (T == 128) return resident_launch<128, 1, 8>(a, G, smem, s, occ_out);
  if (T == 512) {
    if (NCH == 3) return resident_launch<512, 3, 2>(a, G, smem, s, occ_out);
    return resident_launch<512, 4, 2>(a, G, smem, s, occ_out);
  }
  switch (NCH) {
    case 1: return resident_launch<256, 1, 8>(a, G, smem, s, occ_out);
    case 2: return resident_launch<256, 2, 4>(a, G, smem, s, occ_out);
    case 3: return resident_launch<256, 3, 2>(a, G, smem, s, occ_out);
    case 4: return resident_launch<256, 4, 2>(a, G, smem, s, occ_out);
    case 5: return resident_launch<256, 5, 1>(a, G, smem, s, occ_out);
    case 6: return resident_launch<256, 6, 1>(a, G, smem, s, occ_out);
    case 7: return resident_launch<256, 7, 1>(a, G, smem, s, occ_out);
    default: return resident_launch<256, 8, 1>(a, G, smem, s, occ_out);
  }
}

static bool resident_pick(int n, int m, ResCfg* c) {
  if (n < 1 || m < 4 || (m & 3) || m > kResMaxM) return false;
  // rows wider than 4096 columns: one 512-thread CTA per SM keeps two rows per block barrier (16 columns per
  // thread) where two 256-thread CTAs would be down to one (B200OT_RES_WIDE=0 restores the latter)
  const char* ew = getenv("B200OT_RES_WIDE");
  const bool wide = m > 4096 && !(ew && ew[0] == '0');
  const int T = m <= 512 ? 128 : wide ? 512 : 256;
  const int NCH = (m + 4 * T - 1) / (4 * T);
  const int R = NCH == 1 ? 8 : NCH == 2 ? 4 : NCH <= 4 ? 2 : 1;
  const size_t smem_cap = T > 256 ? kResSmemMaxWide : kResSmemMax;
  const int ngroups = (n + R - 1) / R;
  const size_t stage = (size_t)R * T * 4 * NCH * sizeof(float);
  auto fixed_for = [&](int rows_cap) {
    return (size_t)(kResMaxStages * 8 + 256 + 32 * 8 + kResMaxStages * 4 + 2 * R * (T / 32) * 4) + (size_t)T * 4 +
           3 * (size_t)rows_cap * 4 + 128;
  };
  // CTAs per SM the hardware will co-schedule for this instantiation (cached): the cooperative grid may not exceed it
  static int occ_cache[3][9];
  int& occ = occ_cache[T == 128 ? 0 : T == 256 ? 1 : 2][NCH];
  if (occ == 0) {
    int v = 0;
    if (resident_dispatch(T, NCH, nullptr, 0, smem_cap, nullptr, &v) != cudaSuccess) {
      (void)cudaGetLastError();
      v = -1;
    }
    occ = v > 0 ? (v > 2 ? 2 : v) : -1;
  }
  if (occ < 1) return false;
  int G = sm_count() * occ;
  if (G < 1) return false;
  if (G > kResMaxCtas) G = kResMaxCtas;
  if (G > ngroups) G = ngroups;
  const int cnt_max = (ngroups + G - 1) / G;
  const int rows_cap = cnt_max * R;
  const size_t fixed = fixed_for(rows_cap);
  if (fixed + 2 * stage > smem_cap) return false;
  int NG = (int)((smem_cap - fixed) / stage);
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

// auto rule: iterations short enough that launches and the finalize kernel dominate the sweep
static bool resident_wanted(int n, int m) {
  const char* e = getenv("B200OT_RESIDENT");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return (double)n * (double)m <= 8192.0 * 8192.0;
}

static bool g_resident_broken = false;  // a failed cooperative launch disables the path for the process

int resident_try_enqueue(const float* C, int ldc, int n, int m, int iters, const WsPtrs& w, cudaStream_t s) {
  if (iters < 1 || g_resident_broken || !resident_wanted(n, m) || !w.res_ll) return 1;
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
  a.part64 = reinterpret_cast<unsigned long long*>(w.res_ll);
  a.stride = w.m_pad;
  a.g64 = a.part64 + (size_t)kResMaxCtas * w.m_pad;
  a.err64 = a.g64 + w.m_pad;
  a.err_hist = w.err_hist;
  a.ng = c.NG;
  a.rows_cap = c.rows_cap;
  const char* es = getenv("B200OT_RES_SNAKE");
  a.snake = (es && es[0] == '0') ? 0 : 1;
  const char* ee = getenv("B200OT_RES_EVICT");
  a.evict_first = (ee && ee[0] == '1') ? 1 : 0;
  a.red_groups = c.red_groups;
  const char* ens = getenv("B200OT_RES_SPIN_NS");
  a.spin_ns = ens ? atoi(ens) : 20;
  if (a.spin_ns < 0) a.spin_ns = 0;
  for (int left = iters; left > 0;) {
    a.iters = left > kResMaxItersPerLaunch ? kResMaxItersPerLaunch : left;
    left -= a.iters;
    const cudaError_t e = resident_dispatch(c.T, c.NCH, &a, c.G, c.smem, s, nullptr);
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "resident_kernel launch");
      (void)cudaGetLastError();
      g_resident_broken = true;
      // nothing was queued by a failed first launch: the caller can still use the launch-per-sweep path
      return left + a.iters == iters ? 1 : B200OT_E_LAUNCH;
    }
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
           "<=%d groups/CTA, ring %d x %zu B (%s), tagged-word exchange (no grid barrier), smem %zu B/CTA",
           c.G, c.T, 4 * c.NCH, c.R, c.cnt_max, c.NG, (size_t)c.R * c.T * 4 * c.NCH * 4,
           c.cnt_max <= c.NG ? "C shared-memory resident" : (snake ? "snake sweeps" : "forward sweeps"), c.smem);
  return true;
}

}  // namespace b200ot
