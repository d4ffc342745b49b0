// Log-domain Sinkhorn on a materialised fp32 cost matrix (streaming path).
//
// Replaces the reference's ot.sinkhorn / sinkhorn_scaling / ott linear.solve calls
// (MRI_PET_OT_nojax.py:143, perturbot/perturbot/match/utils.py:6-115,
// perturbot/perturbot/match/fot.py:129-134).  One reference iteration is
//     g <- eps log b - eps LSE_i((f_i - C_ij)/eps)      (POT: v = b / K^T u)
//     f <- eps log a - eps LSE_j((g_j - C_ij)/eps)      (POT: u = a / K v)
// followed every check_every iterations by the column-marginal error of
// exp((f+g-C)/eps).  Internally potentials are kept scaled, fs = f*log2(e)/eps, so the
// plan entry is 2^(fs_i + gs_j - C_ij*k), k = log2(e)/eps.
//
// FUSED path (one HBM read of C per iteration).  For sweep k the kernel holds a group of
// R rows on chip: t_ij = 2^(fs_i^{k-1} + gs_j^k - k C_ij) is evaluated once per element
// and kept in registers, the row sums r_i = sum_j t_ij are reduced warp -> CTA -> cluster
// (DSMEM), w_i = a_i / r_i gives fs_i^k = fs_i^{k-1} + log2 w_i, and the same registers
// are folded into per-thread column accumulators s_j += t_ij w_i.  s is exactly the column
// marginal of the plan after iteration k, i.e. both the convergence check and the input of
// the next g update, gs_j^{k+1} = gs_j^k + log2 b_j - log2 s_j.  Every t_ij <= b_j <= 1
// because the preceding g update normalised the columns, so no running max is needed; a
// row or column sum that vanishes (extreme eps in the first iterations) raises the `bad`
// flag and the host replays the chunk on the ROBUST path.
//
// ROBUST path: row pass and column pass as separate sweeps with running-max logsumexp.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "solver_state.cuh"

namespace b200ot {

// =============================================================================
// state / vector initialisation
// =============================================================================
__global__ void init_state_kernel(State* st, b200ot_params prm) {
  State s;
  memset(&s, 0, sizeof(s));
  s.kscale = kLog2e / prm.eps;
  s.eps = prm.eps;
  s.tol = prm.tol;
  s.max_iter = prm.max_iter;
  s.check_every = prm.check_every > 0 ? prm.check_every : 1;
  s.check_phase = prm.check_phase;
  s.err_norm = prm.err_norm;
  s.stop_inclusive = prm.stop_inclusive;
  s.path = prm.path;
  s.err = INFINITY;
  s.initialised = 1;
  s.done = prm.max_iter <= 0 ? 1 : 0;
  s.floor_patience = prm.floor_patience;
  s.best_err = INFINITY;
  s.fs_lo[0] = INFINITY;
  s.fs_hi[0] = -INFINITY;
  s.fs_lo[1] = INFINITY;
  s.fs_hi[1] = -INFINITY;
  s.snap_it = -1;  // no snapshot yet: a failure of the first g update cannot be rewound (rewind_kernel leaves it)
  *st = s;
}

__global__ void __launch_bounds__(256) init_kernel(State* st, float eps, int n, int m,
                                                   const float* __restrict__ a, const float* __restrict__ b,
                                                   const float* __restrict__ f0, const float* __restrict__ g0,
                                                   float* fs, float* gs0, float* gs1, float* wa, float* wb,
                                                   float* log2b) {
  const float k = kLog2e / eps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo = INFINITY, hi = -INFINITY;
  if (i < n) {
    const float v = f0 ? f0[i] * k : 0.f;
    fs[i] = v;
    wa[i] = a[i];
    if (fabsf(v) < INFINITY) lo = hi = v;
  }
  float bl1 = 0.f, bl2 = 0.f;
  if (i < m) {
    const float g = g0 ? g0[i] * k : 0.f;
    gs0[i] = g;
    gs1[i] = g;
    wb[i] = b[i];
    log2b[i] = log2f(b[i]);
    bl1 = fabsf(b[i]);
    bl2 = b[i] * b[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  bl1 = warp_sum(bl1);
  bl2 = warp_sum(bl2);
  if ((threadIdx.x & 31) == 0) {
    if (lo <= hi) {
      atomic_min_float(&st->fs_lo[0], lo);
      atomic_max_float(&st->fs_hi[0], hi);
    }
    if (bl1 > 0.f) {  // |b| only scales the floor guard: summation order is irrelevant
      atomicAdd(&st->b_l1, bl1);
      atomicAdd(&st->b_l2sq, bl2);
    }
  }
}

// =============================================================================
// finalize: column sums -> marginal error -> stopping rule -> next g
// =============================================================================
// part_sum[p][j] (and optional part_max[p][j]: the sums are relative to 2^max) for p < np.
// Writes gs_next = gs_cur + log2 b - log2 s and lets the LAST block (ticket) fold the
// per-block error partials in fixed order and advance the state machine.
// Tagged 64-bit words {fp32 value, 32-bit tag} written by peer GPUs over NVLink (row-sharded solve without NCCL in
// the loop, see reduce_push_kernel): a 64-bit access is single-copy atomic, so the tag of exchange x carries the
// value of exchange x.  Exchange index x = 0 for the prologue, it + 1 for the iteration that follows `it` completed
// ones; slabs are double-buffered by the parity of x.
// Bit 31 of the tag marks a LOG-DOMAIN word: the value is log2 of the rank's column sum instead of the sum.  The
// first g update and the robust two-sweep iterations use it, because their column sums (unbalanced potentials,
// extreme eps) can lie far outside the fp32 range while their logarithms are ordinary numbers.
constexpr unsigned kPeerLogBit = 0x80000000u;
__device__ __forceinline__ unsigned peer_tag(unsigned epoch, int x) {
  return ((epoch & 0x7ffu) << 20) | ((unsigned)(x + 1) & 0xfffffu);
}
// A peer that never pushes (a rank that died, a host stalled for seconds) must not hang or kill this rank:
// after ~4 s the poll gives up and returns NaN, which poisons the error partial, so the state machine stops the
// solve with bad = 1 and the host reads B200OT_E_NUMERIC instead of a dead CUDA context.
__device__ __forceinline__ float peer_poll(const unsigned long long* p, unsigned tag, bool* is_log = nullptr) {
  unsigned long long v;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (((unsigned)(v >> 32) & ~kPeerLogBit) == tag) {
      if (is_log) *is_log = ((unsigned)(v >> 32) & kPeerLogBit) != 0;
      break;
    }
    if (t0 == 0)
      t0 = clock64();
    else if (clock64() - t0 > 8000000000ll)
      return __int_as_float(0x7fc00000);
    __nanosleep(40);
  }
  return __uint_as_float((unsigned)v);
}

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  unsigned long long* buf[kMaxPeers];
};

// One column of the finalize step: marginal-error term and the next g from the column sum s of the new plan.
__device__ __forceinline__ double finalize_column(float s, float l2s, float bj, float log2bj, float gcur_j, int norm,
                                                  float* gnext_j, int* bad) {
  const double d = (double)s - (double)bj;
  const float gn = bj > 0.f ? gcur_j + (log2bj - l2s) : -INFINITY;  // b_j = 0: no mass, v_j = 0
  *gnext_j = gn;
  if (bj > 0.f && !(fabsf(gn) < INFINITY)) *bad = 1;
  return (norm == B200OT_NORM_L1) ? fabs(d) : d * d;
}

__global__ void __launch_bounds__(kFinalizeThreads)
    finalize_kernel(State* st, const float* __restrict__ part_sum, const float* __restrict__ part_max,
                    int np, size_t stride, int m, const float* __restrict__ b,
                    const float* __restrict__ log2b, float* gs0, float* gs1, double* errpart,
                    float* err_hist, int is_prologue, int groups,
                    const unsigned long long* tagged = nullptr, unsigned epoch = 0) {
  if (st->done) return;
  unsigned tag = 0;
  if (tagged) {  // np = world slabs of this exchange's parity
    const int x = is_prologue ? 0 : st->it + 1;
    tag = peer_tag(epoch, x);
    tagged += (size_t)(x & 1) * np * stride;
  }
  const int cur = st->cur;
  const float* gcur = cur ? gs1 : gs0;
  float* gnext = cur ? gs0 : gs1;
  const int norm = st->err_norm;
  // A block covers CB = 256/groups columns; thread (grp, c) folds the partial slabs p = grp, grp+groups, ...
  // so many slabs over few columns still fill the GPU; the `groups` results are combined in fixed order.
  const int CB = kFinalizeThreads / groups;
  const int grp = threadIdx.x / CB, c = threadIdx.x - grp * CB;
  const int j = blockIdx.x * CB + c;
  __shared__ float sh_s[kFinalizeThreads], sh_m[kFinalizeThreads];
  {
    float acc = 0.f, M = -INFINITY;
    if (j < m) {
      if (part_max) {
        for (int p = grp; p < np; p += groups) M = fmaxf(M, part_max[(size_t)p * stride + j]);
        for (int p = grp; p < np; p += groups) {
          const float pm = part_max[(size_t)p * stride + j];
          if (pm > -INFINITY) acc += part_sum[(size_t)p * stride + j] * exp2f(pm - M);
        }
      } else if (tagged) {  // groups == 1: one thread gathers the world's words of its column, rank order
        float v[kMaxPeers];
        bool any_log = false, any_nan = false;
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p) {
          v[p] = -INFINITY;
          if (p < np) {
            bool lg = false;
            v[p] = peer_poll(tagged + (size_t)p * stride + j, tag, &lg);
            any_log |= lg;
            any_nan |= (v[p] != v[p]);  // a rank that lost a sum, or never answered
          }
        }
        if (any_nan) {
          acc = M = __int_as_float(0x7fc00000);
        } else if (any_log) {  // values are log2 of the ranks' sums: log-sum-exp over the ranks, rank order
          float Mx = -INFINITY;
#pragma unroll
          for (int p = 0; p < kMaxPeers; ++p) Mx = fmaxf(Mx, v[p]);
#pragma unroll
          for (int p = 0; p < kMaxPeers; ++p)
            if (v[p] > -INFINITY) acc += exp2f(v[p] - Mx);
          M = Mx > -INFINITY ? Mx + log2f(acc) : -INFINITY;
          acc = exp2f(M);
        } else {
#pragma unroll
          for (int p = 0; p < kMaxPeers; ++p)
            if (p < np) acc += v[p];
          M = log2f(acc);
        }
      } else {
        for (int p = grp; p < np; p += groups) acc += part_sum[(size_t)p * stride + j];
      }
    }
    sh_s[threadIdx.x] = acc;
    sh_m[threadIdx.x] = M;
  }
  __syncthreads();
  double e = 0.0;
  int bad = 0;
  if (j < m && grp == 0) {
    float s, l2s;
    if (tagged) {
      s = sh_s[c];
      l2s = sh_m[c];
    } else if (part_max) {
      float M = -INFINITY;
      for (int g2 = 0; g2 < groups; ++g2) M = fmaxf(M, sh_m[g2 * CB + c]);
      float acc = 0.f;
      for (int g2 = 0; g2 < groups; ++g2) {
        const float pm = sh_m[g2 * CB + c];
        if (pm > -INFINITY) acc += sh_s[g2 * CB + c] * exp2f(pm - M);
      }
      l2s = M + log2f(acc);
      s = exp2f(l2s);
    } else {
      float acc = 0.f;
      for (int g2 = 0; g2 < groups; ++g2) acc += sh_s[g2 * CB + c];
      s = acc;
      l2s = log2f(acc);
    }
    const float bj = b[j];
    const double d = (double)s - (double)bj;
    e = (norm == B200OT_NORM_L1) ? fabs(d) : d * d;
    const float gn = bj > 0.f ? gcur[j] + (log2b[j] - l2s) : -INFINITY;  // b_j = 0: no mass, v_j = 0
    gnext[j] = gn;
    if (bj > 0.f && !(fabsf(gn) < INFINITY)) bad = 1;
  }
  // block reduce (fixed tree)
  __shared__ double sh[kFinalizeThreads / 32];
  __shared__ int shbad;
  if (threadIdx.x == 0) shbad = 0;
  __syncthreads();
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
  if (bad) shbad = 1;
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < kFinalizeThreads / 32; ++w) tot += sh[w];
    errpart[blockIdx.x] = tot;
    if (shbad) atomicExch(&st->bad, 1);
    __threadfence();
    const int t = atomicAdd(&st->ticket, 1);
    is_last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // last block: fold the per-block error partials with all threads (fixed assignment + fixed tree)
  __threadfence();
  {
    double part = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kFinalizeThreads) part += ((volatile double*)errpart)[i];
    part = warp_sum(part);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  for (int w = 0; w < kFinalizeThreads / 32; ++w) tot += sh[w];
  float err = (norm == B200OT_NORM_L2) ? (float)sqrt(tot) : (float)tot;
  st->ticket = 0;
  const int isbad = ((volatile int*)&st->bad)[0];
  if (isbad) {  // fast path lost a sum: stop here, the host rewinds and replays robustly
    st->done = 1;
    return;
  }
  if (is_prologue) {
    st->cur = cur ^ 1;
    return;
  }
  State ls = *st;
  advance_state(ls, err, part_max != nullptr, err_hist);
  *st = ls;
}

// =============================================================================
// FUSED single-sweep kernel
// =============================================================================
struct SweepArgs {
  const float* C;
  long long ldc;
  int n, m;
  State* st;
  float* fs;
  const float* gs0;
  const float* gs1;
  const float* a;
  float* part;  // [nclusters][stride]
  size_t stride;
  int ng;          // ring depth in row groups
  int evict_first; // stream C through L2 with an evict-first policy
  int wq;          // columns per CTA (multiple of 4, <= 2048*NCH): an even split of m over the cluster
  int mode;        // 0 = normal; 1 = stream only (diagnostic: TMA ring without the arithmetic)
  // fused iteration (sweep_lite_kernel only): after the sweep the SAME launch folds the cluster partials, exchanges
  // the column sums with the peer ranks (world > 1), applies the finalize step and advances the state machine
  int stagger_ns;  // tuning (B200OT_STAGGER): clusters with odd id start this many ns late
  int fuse;
  int iters;       // fused form: iterations this launch runs (stops earlier when the stopping rule fires)
  float* gs0w;
  float* gs1w;
  const float* b;
  const float* log2b;
  double* errpart;
  float* err_hist;
  PeerPtrs peers;
  int world, rank;
  size_t xstride;
  unsigned epoch;
};

constexpr int kSweepThreads = 512;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kMaxCluster = 8;
constexpr int kXBuf = 4;  // exchange buffers: a peer can run at most 3 groups ahead of a CTA (see below)

// Layout: a cluster of Q CTAs owns whole rows; CTA q owns columns [q*W, (q+1)*W), W = 512*4*NCH, and
// thread t owns NCH quads of them (quad c = columns c*2048 + 4t ..+3), so every shared-memory read is a
// conflict-free 16-byte access.  Rows arrive in groups of R through a ring of TMA bulk copies
// (cp.async.bulk -> mbarrier complete_tx).  Per group:
//   P1   t_ij = 2^(fs_i + gs_j - k C_ij) into registers, row partial sums warp -> CTA
//   SEND CTA partials to every peer of the cluster with st.async (DSMEM write that completes a
//        transaction on the peer's mbarrier: no cluster barrier, no fence)
//   P2   wait for the Q partials, w_i = a_i / r_i, fs_i += log2 w_i, acc_j += t_ij w_i
// The loop is software-pipelined with two register sets: P1+SEND of group i+1 are issued before P2 of
// group i, so the DSMEM round trip is hidden behind the next group's exponentials.
// Exchange-buffer safety: a peer issues SEND(i+4) only after its P2(i+2), which needs this CTA's
// SEND(i+2), which this CTA issues after its own P2(i); hence 4 buffers (i mod 4) never collide.
template <int NCH, int R>
__global__ void __launch_bounds__(kSweepThreads, 1) sweep_fused_kernel(const SweepArgs p) {
  constexpr int CPT = 4 * NCH;
  constexpr int W = kSweepThreads * CPT;  // columns owned by one CTA
  extern __shared__ __align__(128) unsigned char smem[];

  State* st = p.st;
  if (st->done) return;  // grid-uniform: every CTA of every cluster leaves together
  const int cur = st->cur;
  const float k = st->kscale;
  const float* __restrict__ gs = cur ? p.gs1 : p.gs0;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank();
  const int Q = (int)cluster_nctarank();
  const int cid = (int)cluster_id_x();
  const int NC = (int)cluster_nid_x();
  const int NG = p.ng;

  const long long col0 = (long long)q * p.wq;
  int mvalid = p.m - (int)col0;
  mvalid = mvalid < 0 ? 0 : (mvalid > p.wq ? p.wq : mvalid);

  // Shift of the exponent.  t_ij = 2^(shift + gs_j - k C_ij) must stay in fp32 range; the exact choice
  // cancels in w_i = a_i / r_i.  If the row potentials of the previous iteration span less than 2^48 a
  // single shift (mid-range) is folded into gs once per sweep and saves one FADD per element;
  // otherwise each row uses its own previous potential.
  const int it0 = st->it;
  const float flo = st->fs_lo[it0 & 1], fhi = st->fs_hi[it0 & 1];
  const bool uniform = flo <= fhi && (fhi - flo) < 48.f;  // false when the range is unknown (lo > hi) or infinite
  const float sigma = uniform ? 0.5f * (flo + fhi) : 0.f;

  float* stage = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NG * R * W * sizeof(float));  // [8]
  uint64_t* xbar = full + 8;                                                               // [kXBuf]
  float* red = reinterpret_cast<float*>(xbar + kXBuf);  // [2][kSweepWarps][R]
  float* xch = red + 2 * kSweepWarps * R;                // [kXBuf][R][kMaxCluster]

  float gsv[CPT], acc[CPT];
  bool cvalid[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * (kSweepThreads * 4) + tid * 4;
    cvalid[c] = col < mvalid;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cvalid[c]) g4 = *reinterpret_cast<const float4*>(gs + col0 + col);
    gsv[c * 4 + 0] = g4.x + sigma;
    gsv[c * 4 + 1] = g4.y + sigma;
    gsv[c * 4 + 2] = g4.z + sigma;
    gsv[c * 4 + 3] = g4.w + sigma;
    acc[c * 4 + 0] = acc[c * 4 + 1] = acc[c * 4 + 2] = acc[c * 4 + 3] = 0.f;
  }

  if (tid == 0) {
    for (int s = 0; s < NG; ++s) mbar_init(smem_u32(full + s), 1);
    for (int s = 0; s < kXBuf; ++s) mbar_init(smem_u32(xbar + s), 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_arrive();  // every CTA's barriers exist before any peer sends to them
  cluster_wait();

  const int ngroups = (p.n + R - 1) / R;
  const int cnt = cid < ngroups ? (ngroups - cid + NC - 1) / NC : 0;
  const uint64_t pol = p.evict_first ? policy_evict_first() : 0ull;

  auto issue = [&](int i) {
    const int row0 = (cid + i * NC) * R;
    const int s = i % NG;
    const uint32_t bar = smem_u32(full + s);
    const uint32_t row_bytes = (uint32_t)mvalid * 4u;
    mbar_arrive_expect_tx(bar, row_bytes * R);
    if (row_bytes) {
      for (int r = 0; r < R; ++r) {
        int row = row0 + r;
        row = row < p.n ? row : p.n - 1;  // ragged last group: re-read the last row, its weight is zero
        const float* src = p.C + (long long)row * p.ldc + col0;
        const uint32_t dst = smem_u32(stage + ((size_t)s * R + r) * W);
        if (p.evict_first)
          bulk_g2s_hint(dst, src, row_bytes, bar, pol);
        else
          bulk_g2s(dst, src, row_bytes, bar);
      }
    }
  };

  if (tid == 0) {
    const int pre = cnt < NG ? cnt : NG;
    for (int i = 0; i < pre; ++i) issue(i);
  }

  struct Ctx {
    float sh[R], ar[R];  // sh: exponent shift of the row (sigma or its previous potential)
    int row0, rows;
  };
  float f_lo = INFINITY, f_hi = -INFINITY;  // range of the potentials this thread writes

  // Column validity: the host guarantees mvalid > 2048*(NCH-1) for every CTA, so quads 0..NCH-2 are
  // valid for all threads; only the last quad is masked (warp-uniform skip where a whole warp is out).
  const bool last_ok = cvalid[NCH - 1];
  const bool last_any = __any_sync(0xffffffffu, last_ok);
  const bool stream_only = (p.mode & 1) != 0;  // diagnostic: skip the arithmetic
  const bool no_exchange = (p.mode & 2) != 0;  // diagnostic: skip the DSMEM exchange (wrong row sums)
  const bool no_exp = (p.mode & 4) != 0;       // diagnostic: replace ex2 by a multiply (wrong values)

  // P1 + SEND of group i.  Rows beyond n are loaded as copies of row n-1 (finite data) and get w = 0.
  auto front = [&](int i, float (&t)[R][CPT], Ctx& cx, auto uni_tag) {
    constexpr bool UNI = decltype(uni_tag)::value;
    const int s = i % NG;
    const uint32_t ph = (uint32_t)((i / NG) & 1);
    cx.row0 = (cid + i * NC) * R;
    cx.rows = p.n - cx.row0;
    cx.rows = cx.rows > R ? R : cx.rows;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cx.sh[r] = sigma;
      cx.ar[r] = 0.f;
      if (r < cx.rows) {
        if (!UNI) cx.sh[r] = p.fs[cx.row0 + r];
        cx.ar[r] = p.a[cx.row0 + r];
      }
    }
    mbar_wait(smem_u32(full + s), ph);
    float ps[R];
    auto quad = [&](int r, int c, const float* srow) {
      const float4 v = *reinterpret_cast<const float4*>(srow + c * (kSweepThreads * 4));
      float e0, e1, e2, e3;
      const float2 mk2 = make_float2(-k, -k);
      float2 g0 = make_float2(gsv[c * 4 + 0], gsv[c * 4 + 1]);
      float2 g1 = make_float2(gsv[c * 4 + 2], gsv[c * 4 + 3]);
      if (!UNI) {
        const float2 sh2 = make_float2(cx.sh[r], cx.sh[r]);
        g0 = __fadd2_rn(g0, sh2);
        g1 = __fadd2_rn(g1, sh2);
      }
      const float2 x0 = __ffma2_rn(make_float2(v.x, v.y), mk2, g0);
      const float2 x1 = __ffma2_rn(make_float2(v.z, v.w), mk2, g1);
      if (no_exp) {
        e0 = x0.x * 1e-3f;
        e1 = x0.y * 1e-3f;
        e2 = x1.x * 1e-3f;
        e3 = x1.y * 1e-3f;
      } else {
        e0 = ex2_approx(x0.x);
        e1 = ex2_approx(x0.y);
        e2 = ex2_approx(x1.x);
        e3 = ex2_approx(x1.y);
      }
      t[r][c * 4 + 0] = e0;
      t[r][c * 4 + 1] = e1;
      t[r][c * 4 + 2] = e2;
      t[r][c * 4 + 3] = e3;
    };
    if (!stream_only) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* srow = stage + ((size_t)s * R + r) * W + tid * 4;
#pragma unroll
        for (int c = 0; c < NCH - 1; ++c) quad(r, c, srow);  // straight-line: loads batch ahead of the math
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* srow = stage + ((size_t)s * R + r) * W + tid * 4;
        if (last_any) {
          quad(r, NCH - 1, srow);  // smem beyond mvalid holds stale but finite-or-masked data
          if (!last_ok) t[r][CPT - 4] = t[r][CPT - 3] = t[r][CPT - 2] = t[r][CPT - 1] = 0.f;
        } else {
          t[r][CPT - 4] = t[r][CPT - 3] = t[r][CPT - 2] = t[r][CPT - 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          t[r][c * 4 + 0] = t[r][c * 4 + 1] = t[r][c * 4 + 2] = t[r][c * 4 + 3] = cvalid[c] ? 1.f : 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CPT / 2; ++c) s2 = __fadd2_rn(s2, make_float2(t[r][2 * c], t[r][2 * c + 1]));
      ps[r] = s2.x + s2.y;
    }
    const int par = i & 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float v = warp_sum(ps[r]);
      if (lane == 0) red[(par * kSweepWarps + warp) * R + r] = v;
    }
    __syncthreads();  // every thread has drained ring stage s; red[par] is complete
    const int xb = i % kXBuf;
    if (tid == 0) {
      if (i + NG < cnt) {
        fence_proxy_async();
        issue(i + NG);
      }
      if (!no_exchange) mbar_arrive_expect_tx(smem_u32(xbar + xb), (uint32_t)(Q * R * 4));
    }
    if (tid < R * Q && !no_exchange) {
      const int r = tid / Q, qq = tid - r * Q;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kSweepWarps; ++w) v += red[(par * kSweepWarps + w) * R + r];
      const uint32_t dst = map_to_cta(smem_u32(xch + (xb * R + r) * kMaxCluster + q), (uint32_t)qq);
      const uint32_t bar = map_to_cta(smem_u32(xbar + xb), (uint32_t)qq);
      st_async_f32(dst, v, bar);
    }
  };

  // P2 of group i
  auto back = [&](int i, float (&t)[R][CPT], const Ctx& cx) {
    const int xb = i % kXBuf;
    if (!no_exchange) mbar_wait(smem_u32(xbar + xb), (uint32_t)((i / kXBuf) & 1));
    float wr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float rt = 0.f;
      if (no_exchange) {
        rt = red[((i & 1) * kSweepWarps) * R + r] * (float)(Q * kSweepWarps);
      } else {
#pragma unroll
        for (int qq = 0; qq < kMaxCluster; ++qq)
          if (qq < Q) rt += xch[(xb * R + r) * kMaxCluster + qq];
      }
      const bool live = (r < cx.rows) && (cx.ar[r] > 0.f);
      wr[r] = live ? __fdividef(cx.ar[r], rt) : 0.f;
      if (q == 0 && tid == r && r < cx.rows && !stream_only) {
        const float fnew = live ? cx.sh[r] + (log2f(cx.ar[r]) - log2f(rt)) : -INFINITY;
        p.fs[cx.row0 + r] = fnew;
        if (live) {
          if (fabsf(fnew) < INFINITY) {
            f_lo = fminf(f_lo, fnew);
            f_hi = fmaxf(f_hi, fnew);
          } else {
            atomicExch(&st->bad, 1);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 w2 = make_float2(wr[r], wr[r]);
#pragma unroll
      for (int c = 0; c < CPT / 2; ++c) {
        const float2 a2 = __ffma2_rn(make_float2(t[r][2 * c], t[r][2 * c + 1]), w2, make_float2(acc[2 * c], acc[2 * c + 1]));
        acc[2 * c] = a2.x;
        acc[2 * c + 1] = a2.y;
      }
    }
  };

  float tA[R][CPT], tB[R][CPT];
  Ctx cA, cB;
  // F0 F1 B0 F2 B1 F3 B2 ... : even groups use register set A, odd groups set B
  auto run = [&](auto uni_tag) {
    for (int i = -1; i < cnt; i += 2) {
      if (i + 1 < cnt) front(i + 1, tA, cA, uni_tag);
      if (i >= 0) back(i, tB, cB);
      if (i + 2 < cnt) front(i + 2, tB, cB, uni_tag);
      if (i + 1 < cnt) back(i + 1, tA, cA);
    }
  };
  if (uniform)
    run(std::true_type{});
  else
    run(std::false_type{});

  if (q == 0 && tid < R && f_lo <= f_hi) {
    atomic_min_float(&st->fs_lo[(it0 + 1) & 1], f_lo);
    atomic_max_float(&st->fs_hi[(it0 + 1) & 1], f_hi);
  }

#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (cvalid[c]) {
      const int col = c * (kSweepThreads * 4) + tid * 4;
      *reinterpret_cast<float4*>(p.part + (size_t)cid * p.stride + col0 + col) =
          make_float4(acc[c * 4 + 0], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
    }
  }
  cluster_arrive();  // no CTA retires while a peer may still address its shared memory
  cluster_wait();
}

constexpr int kLiteThreads = 256;
constexpr int kLiteWarps = kLiteThreads / 32;
constexpr size_t kLiteSmemMax = 113 * 1024;

// =============================================================================
// Fused iteration tail: fold -> (peer exchange) -> finalize -> state machine, inside the sweep's own launch
// =============================================================================
// Runs after every CTA of the (cooperative, fully co-resident) grid has written its cluster's column partials.
// One grid barrier makes the slabs visible; then column j is owned by thread (j mod G*T): it folds the NC slabs in
// fixed order, (world > 1) stores the rank's sum as a tagged word into slab [parity][rank] of every peer's exchange
// buffer over NVLink and polls the world slabs of its own buffer, adds them in rank order (identical bits on every
// rank), and applies the finalize step.  The last CTA to finish (ticket) folds the error partials in fixed order and
// advances the state machine -- exactly what finalize_kernel does, minus two kernel launches and their gaps.
// `scratch`: >= 128 bytes of the kernel's DYNAMIC shared memory (the TMA ring is idle by now).  Static shared memory
// would add to the 113 KB the kernel is sized for and cost the second CTA per SM.
// Grid barrier of a cooperative (fully co-resident) launch: counter + generation word in the state block.
__device__ __forceinline__ void grid_sync(State* st, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned gen;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&st->gbar_gen) : "memory");
    __threadfence();
    const unsigned t = atomicAdd(&st->gbar_count, 1u);
    if (t == G - 1) {
      st->gbar_count = 0;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&st->gbar_gen), "r"(gen + 1) : "memory");
    } else {
      unsigned now;
      const long long t0 = clock64();
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(&st->gbar_gen) : "memory");
        if (now == gen) {
          __nanosleep(64);
          if (clock64() - t0 > 8000000000ll) {  // a lost CTA must not hang the box: poison the solve instead
            atomicExch(&st->bad, 1);
            atomicExch(&st->done, 1);
            break;
          }
        }
      } while (now == gen);
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __noinline__ void sweep_fused_tail(const SweepArgs& p, State* st, int cur, int it0, int NC,
                                              unsigned char* scratch) {
  const int tid = threadIdx.x;
  const unsigned G = gridDim.x;
  double* sh_e = reinterpret_cast<double*>(scratch);  // [kLiteThreads / 32]
  int& sh_bad = *reinterpret_cast<int*>(scratch + 64);
  int& sh_last = *reinterpret_cast<int*>(scratch + 68);
  if (tid == 0) sh_bad = 0;
  grid_sync(st, G);  // every cluster's partial slab is complete and visible

  const float* gcur = cur ? p.gs1w : p.gs0w;
  float* gnext = cur ? p.gs0w : p.gs1w;
  const int norm = st->err_norm;
  // a row sum that vanished on THIS rank (bad raised by the sweep) must stop every rank in the same iteration:
  // the rank's column sums travel as NaN, every rank's error becomes NaN and all state machines stop with bad = 1
  const bool lbad = ((volatile int*)&st->bad)[0] != 0;
  const int x = it0 + 1;  // exchange index of this iteration
  const unsigned tag = peer_tag(p.epoch, x);
  const unsigned long long word_hi = (unsigned long long)tag << 32;
  const size_t slab0 = (size_t)(x & 1) * p.world * p.xstride;
  double e = 0.0;
  int bad = 0;
  for (int j = (int)blockIdx.x * kLiteThreads + tid; j < p.m; j += (int)G * kLiteThreads) {
    float s = 0.f;
    for (int c = 0; c < NC; ++c) s += __ldcg(p.part + (size_t)c * p.stride + j);  // written by other SMs: L2
    if (lbad) s = __int_as_float(0x7fc00000);
    if (p.world > 1) {
      const unsigned long long word = word_hi | (unsigned long long)__float_as_uint(s);
      const size_t off = slab0 + (size_t)p.rank * p.xstride + j;
      for (int r = 0; r < p.world; ++r) {
        const int dst = (p.rank + r) % p.world;  // everyone starts with itself: spreads the ranks over the links
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.peers.buf[dst] + off), "l"(word) : "memory");
      }
      const unsigned long long* mine = p.peers.buf[p.rank] + slab0 + j;
      s = 0.f;
      for (int r = 0; r < p.world; ++r) s += peer_poll(mine + (size_t)r * p.xstride, tag);
    }
    float gn;
    e += finalize_column(s, log2f(s), p.b[j], p.log2b[j], __ldcg(gcur + j), norm, &gn, &bad);
    gnext[j] = gn;
  }
  e = warp_sum(e);
  if ((tid & 31) == 0) sh_e[tid >> 5] = e;
  if (bad) sh_bad = 1;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < kLiteThreads / 32; ++w) tot += sh_e[w];
    p.errpart[blockIdx.x] = tot;
    if (sh_bad) atomicExch(&st->bad, 1);
    __threadfence();
    const int t = atomicAdd(&st->ticket, 1);
    sh_last = (t == (int)G - 1);
  }
  __syncthreads();
  if (sh_last) {  // the last CTA folds the error partials in fixed order and advances the state machine
    __threadfence();
    double part = 0.0;
    for (unsigned i = tid; i < G; i += kLiteThreads) part += ((volatile double*)p.errpart)[i];
    part = warp_sum(part);
    __syncthreads();
    if ((tid & 31) == 0) sh_e[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kLiteThreads / 32; ++w) tot += sh_e[w];
      // NaN error (a vanished column sum, a peer that never answered): stop with bad = 1
      const float err = (norm == B200OT_NORM_L2) ? (float)sqrt(tot) : (float)tot;
      st->ticket = 0;
      if (!(tot == tot)) atomicExch(&st->bad, 1);
      __threadfence();
      if (((volatile int*)&st->bad)[0]) {
        st->done = 1;
      } else {
        State ls = *st;
        ls.gbar_count = ((volatile State*)st)->gbar_count;  // the barrier words belong to grid_sync, not to this copy
        ls.gbar_gen = ((volatile State*)st)->gbar_gen;
        advance_state(ls, err, false, p.err_hist);
        // field-wise write-back of what advance_state may change (never the barrier words, which other CTAs
        // are updating concurrently)
        st->it = ls.it;
        st->cur = ls.cur;
        st->err = ls.err;
        st->n_err = ls.n_err;
        st->converged = ls.converged;
        st->fs_lo[0] = ls.fs_lo[0];
        st->fs_lo[1] = ls.fs_lo[1];
        st->fs_hi[0] = ls.fs_hi[0];
        st->fs_hi[1] = ls.fs_hi[1];
        st->best_err = ls.best_err;
        st->stall = ls.stall;
        st->floor_hit = ls.floor_hit;
        __threadfence();
        st->done = ls.done;
      }
    }
  }
  grid_sync(st, G);  // the new state (iteration count, current g buffer, done) is visible to every CTA
}

// =============================================================================
// FUSED single-sweep kernel, "lite" form: 256-thread CTAs, two per SM
// =============================================================================
// Same mathematics and data path as sweep_fused_kernel, but latency is hidden by occupancy instead of by
// software pipelining: each CTA keeps ONE register set of exponentials (32 per thread), so two CTAs of
// different clusters share an SM and one computes while the other sits in its reduction / DSMEM exchange.

// The round-1 argument block, kept as it was: the code generated for the plain kernel depends on it.
struct SweepArgsPlain {
  const float* C;
  long long ldc;
  int n, m;
  State* st;
  float* fs;
  const float* gs0;
  const float* gs1;
  const float* a;
  float* part;  // [nclusters][stride]
  size_t stride;
  int ng;          // ring depth in row groups
  int evict_first; // stream C through L2 with an evict-first policy
  int wq;          // columns per CTA
  int mode;        // diagnostics of the pipelined kernel; unused here
};

// Plain form: ONE sweep per launch, finalize is its own launch.  This is the round-1 kernel, kept verbatim: it sits
// at 120 of the 128 registers that two CTAs per SM allow, and a same-box A/B showed that carrying the bookkeeping of
// the persistent form through its row loop (more live state -> rematerialised addresses, +13 % instructions per row)
// costs 7 % of the bandwidth (336 -> 313 it/s at 65536^2).  The persistent fused form below is a separate kernel.
// Also measured against this kernel on one box (profiles/r02_sweep_ab.json; 333 it/s here): the three fp32 streams of
// the row loop as f32x2 instructions (FFMA2 / FADD2: 48 instead of 127 FP32 issue slots per row) 324 it/s, the same
// with the ring slot kept as a running counter instead of i % NG 303 it/s.  The kernel runs under the board's power
// cap (1.73 of 1.965 GHz) and is bound by the per-row dependency chain, not by issue slots.  Diagnostic builds on
// another box (339 it/s): without the cluster exchange (wrong sums) 363 it/s, without the warp reduction either 373
// it/s = 0.98 of the measured HBM peak at 1.67 GHz -- the ceiling of this decomposition under the power cap; folding
// the 256 partials in one warp after the barrier instead of a shuffle tree in every warp: 338 it/s (no change).
template <int NCH>
__global__ void __launch_bounds__(kLiteThreads, 2) sweep_lite_plain_kernel(const SweepArgsPlain p) {
  constexpr int CPT = 4 * NCH;
  constexpr int W = kLiteThreads * CPT;
  extern __shared__ __align__(128) unsigned char smem[];

  State* st = p.st;
  if (st->done) return;
  const int cur = st->cur;
  const float k = st->kscale;
  const float* __restrict__ gs = cur ? p.gs1 : p.gs0;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank();
  const int Q = (int)cluster_nctarank();
  const int cid = (int)cluster_id_x();
  const int NC = (int)cluster_nid_x();
  const int NG = p.ng;

  const long long col0 = (long long)q * p.wq;
  int mvalid = p.m - (int)col0;
  mvalid = mvalid < 0 ? 0 : (mvalid > p.wq ? p.wq : mvalid);

  const int it0 = st->it;
  const float flo = st->fs_lo[it0 & 1], fhi = st->fs_hi[it0 & 1];
  const bool uniform = flo <= fhi && (fhi - flo) < 48.f;
  const float sigma = uniform ? 0.5f * (flo + fhi) : 0.f;

  float* stage = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NG * W * sizeof(float));  // [8]
  uint64_t* xbar = full + 8;                                                            // [kXBuf]
  float* red = reinterpret_cast<float*>(xbar + kXBuf);  // [2][kLiteWarps]
  float* xch = red + 2 * kLiteWarps;                     // [kXBuf][kMaxCluster]

  float gsv[CPT], acc[CPT];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * (kLiteThreads * 4) + tid * 4;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < mvalid) g4 = *reinterpret_cast<const float4*>(gs + col0 + col);
    gsv[c * 4 + 0] = g4.x + sigma;
    gsv[c * 4 + 1] = g4.y + sigma;
    gsv[c * 4 + 2] = g4.z + sigma;
    gsv[c * 4 + 3] = g4.w + sigma;
    acc[c * 4 + 0] = acc[c * 4 + 1] = acc[c * 4 + 2] = acc[c * 4 + 3] = 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < NG; ++s) mbar_init(smem_u32(full + s), 1);
    for (int s = 0; s < kXBuf; ++s) mbar_init(smem_u32(xbar + s), 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  const int cnt = cid < p.n ? (p.n - cid + NC - 1) / NC : 0;  // rows cid, cid + NC, ...
  const uint64_t pol = p.evict_first ? policy_evict_first() : 0ull;
  const uint32_t row_bytes = (uint32_t)mvalid * 4u;
  auto issue = [&](int i) {
    const int s = i % NG;
    const uint32_t bar = smem_u32(full + s);
    mbar_arrive_expect_tx(bar, row_bytes);
    if (row_bytes) {
      const float* src = p.C + (long long)(cid + i * NC) * p.ldc + col0;
      const uint32_t dst = smem_u32(stage + (size_t)s * W);
      if (p.evict_first)
        bulk_g2s_hint(dst, src, row_bytes, bar, pol);
      else
        bulk_g2s(dst, src, row_bytes, bar);
    }
  };
  if (tid == 0) {
    const int pre = cnt < NG ? cnt : NG;
    for (int i = 0; i < pre; ++i) issue(i);
  }

  const bool last_ok = (NCH - 1) * (kLiteThreads * 4) + tid * 4 < mvalid;
  const bool last_any = __any_sync(0xffffffffu, last_ok);
  float f_lo = INFINITY, f_hi = -INFINITY;

  auto row_loop = [&](auto uni_tag) {
    constexpr bool UNI = decltype(uni_tag)::value;
    float a_next = cnt > 0 ? p.a[cid] : 0.f;
    float sh_next = (!UNI && cnt > 0) ? p.fs[cid] : sigma;
    int qsel = 0;  // CTA of the cluster that writes this row's potential
    for (int i = 0; i < cnt; ++i) {
      const int row = cid + i * NC;
      const int s = i % NG;
      const float ar = a_next, sh = sh_next;
      if (i + 1 < cnt) {  // prefetch the next row's scalars so their latency is off the critical path
        a_next = p.a[row + NC];
        if (!UNI) sh_next = p.fs[row + NC];
      }
      mbar_wait(smem_u32(full + s), (uint32_t)((i / NG) & 1));
      float t[CPT];
      const float* srow = stage + (size_t)s * W + tid * 4;
      auto quad = [&](int c) {
        const float4 v = *reinterpret_cast<const float4*>(srow + c * (kLiteThreads * 4));
        if (UNI) {
          t[c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0]));
          t[c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1]));
          t[c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2]));
          t[c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3]));
        } else {
          t[c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0] + sh));
          t[c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1] + sh));
          t[c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2] + sh));
          t[c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3] + sh));
        }
      };
#pragma unroll
      for (int c = 0; c < NCH - 1; ++c) quad(c);
      if (last_any) {
        quad(NCH - 1);
        if (!last_ok) t[CPT - 4] = t[CPT - 3] = t[CPT - 2] = t[CPT - 1] = 0.f;
      } else {
        t[CPT - 4] = t[CPT - 3] = t[CPT - 2] = t[CPT - 1] = 0.f;
      }
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) ps += (t[c * 4 + 0] + t[c * 4 + 1]) + (t[c * 4 + 2] + t[c * 4 + 3]);
      ps = warp_sum(ps);
      const int par = i & 1;
      if (lane == 0) red[par * kLiteWarps + warp] = ps;
      __syncthreads();  // ring stage drained by every warp; red[par] complete
      const int xb = i % kXBuf;
      // per-row serial chores rotate over warps (and the potential update over the CTAs of the cluster) so that
      // no warp is systematically the last one at the next block barrier
      if (lane == 0 && warp == ((i + 4) & (kLiteWarps - 1)) && i + NG < cnt) {
        fence_proxy_async();
        issue(i + NG);
      }
      if (warp == ((i + 6) & (kLiteWarps - 1))) {  // the row-sum exchange rotates too
        if (lane == 0) mbar_arrive_expect_tx(smem_u32(xbar + xb), (uint32_t)(Q * 4));
        if (lane < Q) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kLiteWarps; ++w) v += red[par * kLiteWarps + w];
          st_async_f32(map_to_cta(smem_u32(xch + xb * kMaxCluster + q), (uint32_t)lane), v,
                       map_to_cta(smem_u32(xbar + xb), (uint32_t)lane));
        }
      }
      mbar_wait(smem_u32(xbar + xb), (uint32_t)((i / kXBuf) & 1));
      float rt = 0.f;
#pragma unroll
      for (int qq = 0; qq < kMaxCluster; ++qq)
        if (qq < Q) rt += xch[xb * kMaxCluster + qq];
      const bool live = ar > 0.f;
      const float w = live ? __fdividef(ar, rt) : 0.f;
      if (q == qsel && lane == 0 && warp == ((i + 2) & (kLiteWarps - 1))) {
        const float fnew = live ? sh + (log2f(ar) - log2f(rt)) : -INFINITY;
        p.fs[row] = fnew;
        if (live) {
          if (fabsf(fnew) < INFINITY) {
            f_lo = fminf(f_lo, fnew);
            f_hi = fmaxf(f_hi, fnew);
          } else {
            atomicExch(&st->bad, 1);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[c] = fmaf(t[c], w, acc[c]);
      qsel = (qsel + 1 == Q) ? 0 : qsel + 1;
    }
  };
  if (uniform)
    row_loop(std::true_type{});
  else
    row_loop(std::false_type{});

  if (lane == 0 && f_lo <= f_hi) {  // every thread that wrote potentials
    atomic_min_float(&st->fs_lo[(it0 + 1) & 1], f_lo);
    atomic_max_float(&st->fs_hi[(it0 + 1) & 1], f_hi);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = c * (kLiteThreads * 4) + tid * 4;
    if (col < mvalid)
      *reinterpret_cast<float4*>(p.part + (size_t)cid * p.stride + col0 + col) =
          make_float4(acc[c * 4 + 0], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
  }
  cluster_arrive();
  cluster_wait();
}


// FUSE = false: one sweep per launch (the round-1 kernel; finalize is its own launch).  FUSE = true: the persistent
// form -- up to p.iters iterations per launch, each ending in sweep_fused_tail.  Two instantiations, because the
// bookkeeping of the persistent form costs registers and the plain kernel sits exactly at the 128-register limit of
// two CTAs per SM (same-box A/B: 336 vs 313 it/s when the plain path carried the persistent bookkeeping).
template <int NCH, bool FUSE>
__global__ void __launch_bounds__(kLiteThreads, 2) sweep_lite_kernel(const __grid_constant__ SweepArgs p) {
  constexpr int CPT = 4 * NCH;
  constexpr int W = kLiteThreads * CPT;
  extern __shared__ __align__(128) unsigned char smem[];

  State* st = p.st;
  if (st->done) return;
  const float k = st->kscale;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = (int)cluster_ctarank();
  const int Q = (int)cluster_nctarank();
  const int cid = (int)cluster_id_x();
  const int NC = (int)cluster_nid_x();
  const int NG = p.ng;

  const long long col0 = (long long)q * p.wq;
  int mvalid = p.m - (int)col0;
  mvalid = mvalid < 0 ? 0 : (mvalid > p.wq ? p.wq : mvalid);

  float* stage = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NG * W * sizeof(float));  // [8]
  uint64_t* xbar = full + 8;                                                            // [kXBuf]
  float* red = reinterpret_cast<float*>(xbar + kXBuf);  // [2][kLiteWarps]
  float* xch = red + 2 * kLiteWarps;                     // [kXBuf][kMaxCluster]
  // (128 bytes after xch are the scratch of the fused tail)

  if (tid == 0) {
    for (int s = 0; s < NG; ++s) mbar_init(smem_u32(full + s), 1);
    for (int s = 0; s < kXBuf; ++s) mbar_init(smem_u32(xbar + s), 1);
    fence_mbar_init();
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  // One launch runs `iters` iterations when the fused tail is on (persistent form): the rows of this cluster are
  // the same in every iteration, so the TMA ring simply keeps running across the iteration boundary -- while the
  // grid folds, exchanges and finalizes, the first rows of the next sweep are already on their way.
  const int n_it = FUSE ? (p.iters > 0 ? p.iters : 1) : 1;
  const int cnt = cid < p.n ? (p.n - cid + NC - 1) / NC : 0;  // rows cid, cid + NC, ...
  const int total_rows = cnt * n_it;  // the host keeps this below 2^30
  const uint64_t pol = p.evict_first ? policy_evict_first() : 0ull;
  const uint32_t row_bytes = (uint32_t)mvalid * 4u;
  // The two CTAs that share an SM belong to different clusters.  All clusters start together, so their
  // exponential phases (MUFU-bound) coincide and their exchange phases (SM idle) coincide too; starting every other
  // cluster half a row period late lets one CTA of an SM compute while the other waits for its cluster.
  if (p.stagger_ns > 0 && (cid & 1)) __nanosleep((unsigned)p.stagger_ns);
  // s: ring slot; irow: index of the row in the sweep of this cluster
  auto issue = [&](int s, int irow) {
    const uint32_t bar = smem_u32(full + s);
    mbar_arrive_expect_tx(bar, row_bytes);
    if (row_bytes) {
      const float* src = p.C + (long long)(cid + irow * NC) * p.ldc + col0;
      const uint32_t dst = smem_u32(stage + (size_t)s * W);
      if (p.evict_first)
        bulk_g2s_hint(dst, src, row_bytes, bar, pol);
      else
        bulk_g2s(dst, src, row_bytes, bar);
    }
  };
  if (tid == 0 && cnt > 0) {
    const int pre = total_rows < NG ? total_rows : NG;
    for (int gq = 0; gq < pre; ++gq) issue(gq, gq % cnt);
  }

  const bool last_ok = (NCH - 1) * (kLiteThreads * 4) + tid * 4 < mvalid;
  const bool last_any = __any_sync(0xffffffffu, last_ok);
  // Ring / exchange bookkeeping runs over ALL rows of the launch (the ring does not restart at an iteration
  // boundary) and is kept as incrementing warp-uniform counters: no division by the run-time ring depth per row.
  int slot = 0;                     // ring slot of the row being consumed
  uint32_t sphase = 0;              // its mbarrier phase
  uint32_t rowctr = 0;              // rows consumed: parity of the row-partial buffer (bit 0), exchange buffer
                                    // (bits 0-1, kXBuf = 4 deep) and its mbarrier phase (bit 2)
  int to_issue = total_rows - (total_rows < NG ? total_rows : NG);  // rows still to be issued by the refill
  int inext = cnt > 0 ? NG % cnt : 0;                               // sweep index of the next row to issue

  for (int kit = 0; kit < n_it; ++kit) {
    // ---- per-iteration state (rewritten by the tail of the previous iteration: read it past L1) ----
    const int cur = ((volatile int*)&st->cur)[0];
    const int it0 = ((volatile int*)&st->it)[0];
    const float flo = ((volatile float*)st->fs_lo)[it0 & 1], fhi = ((volatile float*)st->fs_hi)[it0 & 1];
    const bool uniform = flo <= fhi && (fhi - flo) < 48.f;
    const float sigma = uniform ? 0.5f * (flo + fhi) : 0.f;
    const float* gs = cur ? p.gs1 : p.gs0;
    float gsv[CPT], acc[CPT];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * (kLiteThreads * 4) + tid * 4;
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < mvalid) g4 = __ldcg(reinterpret_cast<const float4*>(gs + col0 + col));
      gsv[c * 4 + 0] = g4.x + sigma;
      gsv[c * 4 + 1] = g4.y + sigma;
      gsv[c * 4 + 2] = g4.z + sigma;
      gsv[c * 4 + 3] = g4.w + sigma;
      acc[c * 4 + 0] = acc[c * 4 + 1] = acc[c * 4 + 2] = acc[c * 4 + 3] = 0.f;
    }
    float f_lo = INFINITY, f_hi = -INFINITY;

    auto row_loop = [&](auto uni_tag) {
      constexpr bool UNI = decltype(uni_tag)::value;
      float a_next = cnt > 0 ? p.a[cid] : 0.f;
      float sh_next = (!UNI && cnt > 0) ? __ldcg(p.fs + cid) : sigma;
      int qsel = 0;  // CTA of the cluster that writes this row's potential
      for (int i = 0; i < cnt; ++i) {
        const int row = cid + i * NC;
        const int s = slot;
        const float ar = a_next, sh = sh_next;
        if (i + 1 < cnt) {  // prefetch the next row's scalars so their latency is off the critical path
          a_next = p.a[row + NC];
          if (!UNI) sh_next = __ldcg(p.fs + row + NC);
        }
        mbar_wait(smem_u32(full + s), sphase);
        float t[CPT];
        const float* srow = stage + (size_t)s * W + tid * 4;
        auto quad = [&](int c) {
          const float4 v = *reinterpret_cast<const float4*>(srow + c * (kLiteThreads * 4));
          if (UNI) {
            t[c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0]));
            t[c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1]));
            t[c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2]));
            t[c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3]));
          } else {
            t[c * 4 + 0] = ex2_approx(fmaf(v.x, -k, gsv[c * 4 + 0] + sh));
            t[c * 4 + 1] = ex2_approx(fmaf(v.y, -k, gsv[c * 4 + 1] + sh));
            t[c * 4 + 2] = ex2_approx(fmaf(v.z, -k, gsv[c * 4 + 2] + sh));
            t[c * 4 + 3] = ex2_approx(fmaf(v.w, -k, gsv[c * 4 + 3] + sh));
          }
        };
#pragma unroll
        for (int c = 0; c < NCH - 1; ++c) quad(c);
        if (last_any) {
          quad(NCH - 1);
          if (!last_ok) t[CPT - 4] = t[CPT - 3] = t[CPT - 2] = t[CPT - 1] = 0.f;
        } else {
          t[CPT - 4] = t[CPT - 3] = t[CPT - 2] = t[CPT - 1] = 0.f;
        }
        float ps = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) ps += (t[c * 4 + 0] + t[c * 4 + 1]) + (t[c * 4 + 2] + t[c * 4 + 3]);
        ps = warp_sum(ps);
        const int par = (int)(rowctr & 1u);
        const int xb = (int)(rowctr & (kXBuf - 1));
        if (lane == 0) red[par * kLiteWarps + warp] = ps;
        __syncthreads();  // ring stage drained by every warp; red[par] complete
        // per-row serial chores rotate over warps (and the potential update over the CTAs of the cluster) so that
        // no warp is systematically the last one at the next block barrier
        if (to_issue > 0) {  // the slot just drained takes the row NG ahead -- of the NEXT sweep past the end
          if (lane == 0 && warp == ((i + 4) & (kLiteWarps - 1))) {
            fence_proxy_async();
            issue(s, inext);
          }
          --to_issue;
          if (++inext == cnt) inext = 0;
        }
        if (warp == ((i + 6) & (kLiteWarps - 1))) {  // the row-sum exchange rotates too
          if (lane == 0) mbar_arrive_expect_tx(smem_u32(xbar + xb), (uint32_t)(Q * 4));
          if (lane < Q) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kLiteWarps; ++w) v += red[par * kLiteWarps + w];
            st_async_f32(map_to_cta(smem_u32(xch + xb * kMaxCluster + q), (uint32_t)lane), v,
                         map_to_cta(smem_u32(xbar + xb), (uint32_t)lane));
          }
        }
        mbar_wait(smem_u32(xbar + xb), (rowctr >> 2) & 1u);
        float rt = 0.f;
#pragma unroll
        for (int qq = 0; qq < kMaxCluster; ++qq)
          if (qq < Q) rt += xch[xb * kMaxCluster + qq];
        const bool live = ar > 0.f;
        const float w = live ? __fdividef(ar, rt) : 0.f;
        if (q == qsel && lane == 0 && warp == ((i + 2) & (kLiteWarps - 1))) {
          const float fnew = live ? sh + (log2f(ar) - log2f(rt)) : -INFINITY;
          p.fs[row] = fnew;
          if (live) {
            if (fabsf(fnew) < INFINITY) {
              f_lo = fminf(f_lo, fnew);
              f_hi = fmaxf(f_hi, fnew);
            } else {
              atomicExch(&st->bad, 1);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = fmaf(t[c], w, acc[c]);
        qsel = (qsel + 1 == Q) ? 0 : qsel + 1;
        if (++slot == NG) {
          slot = 0;
          sphase ^= 1u;
        }
        ++rowctr;
      }
    };
    if (uniform)
      row_loop(std::true_type{});
    else
      row_loop(std::false_type{});

    if (lane == 0 && f_lo <= f_hi) {  // every thread that wrote potentials
      atomic_min_float(&st->fs_lo[(it0 + 1) & 1], f_lo);
      atomic_max_float(&st->fs_hi[(it0 + 1) & 1], f_hi);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = c * (kLiteThreads * 4) + tid * 4;
      if (col < mvalid)
        *reinterpret_cast<float4*>(p.part + (size_t)cid * p.stride + col0 + col) =
            make_float4(acc[c * 4 + 0], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
    }
    if constexpr (!FUSE) {
      break;
    } else {
      // (cur and it are re-read here instead of being kept live across the row loop)
      sweep_fused_tail(p, st, ((volatile int*)&st->cur)[0], ((volatile int*)&st->it)[0], NC,
                       smem + (size_t)NG * W * sizeof(float) + 8 * 8 + kXBuf * 8 + (2 * kLiteWarps + kXBuf * kMaxCluster) * 4);
      if (((volatile int*)&st->done)[0]) break;  // grid-uniform: read after the tail's closing grid barrier
    }
  }

  // rows of sweeps that will not run (the stopping rule fired inside the launch) are still in flight: a CTA must
  // not retire with bulk copies outstanding into its shared memory
  if (cnt > 0) {
    // issued so far = total_rows - to_issue; the rows not consumed sit in the slots from `slot` onwards
    const int outstanding = (total_rows - to_issue) - (int)rowctr;  // issued - consumed
    for (int j = 0; j < outstanding; ++j) {
      mbar_wait(smem_u32(full + slot), sphase);
      if (++slot == NG) {
        slot = 0;
        sphase ^= 1u;
      }
    }
  }
  cluster_arrive();
  cluster_wait();
}

// fold the per-cluster column partials of this rank into one vector (row-sharded mode)
__global__ void reduce_parts_kernel(const State* st, const float* __restrict__ part_sum,
                                    const float* __restrict__ part_max, int np, size_t stride, int m,
                                    float* __restrict__ s_out, int accumulate) {
  if (st->done) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  float acc = accumulate ? s_out[j] : 0.f;
  if (part_max) {
    for (int p = 0; p < np; ++p) {
      const float pm = part_max[(size_t)p * stride + j];
      if (pm > -INFINITY) acc += part_sum[(size_t)p * stride + j] * exp2f(pm);
    }
  } else {
    for (int p = 0; p < np; ++p) acc += part_sum[(size_t)p * stride + j];
  }
  s_out[j] = acc;
}

// Row-sharded solve without NCCL in the loop: fold this rank's cluster partials and PUSH the result, as tagged
// words, into slab [parity][rank] of every peer's exchange buffer (peer memory mapped over NVLink / NVSwitch with
// CUDA IPC; stores are posted, nobody waits).  The finalize kernel of each rank then polls its own buffer.
__global__ void __launch_bounds__(256)
    reduce_push_kernel(const State* st, const float* __restrict__ part_sum, const float* __restrict__ part_max, int np,
                       size_t stride, int m, PeerPtrs peers, int world, int rank, size_t xstride, unsigned epoch,
                       int is_prologue) {
  if (st->done) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  float acc = 0.f;
  unsigned flag = 0;
  if (part_max) {  // running-max partials (first g update, robust iterations): send log2 of the rank's sum
    float Mx = -INFINITY;
    for (int p = 0; p < np; ++p) Mx = fmaxf(Mx, part_max[(size_t)p * stride + j]);
    for (int p = 0; p < np; ++p) {
      const float pm = part_max[(size_t)p * stride + j];
      if (pm > -INFINITY) acc += part_sum[(size_t)p * stride + j] * exp2f(pm - Mx);
    }
    acc = Mx > -INFINITY ? Mx + log2f(acc) : -INFINITY;
    flag = kPeerLogBit;
  } else {
    for (int p = 0; p < np; ++p) acc += part_sum[(size_t)p * stride + j];
  }
  if (st->bad) acc = __int_as_float(0x7fc00000);  // a lost sum on this rank stops every rank (NaN error on all of them)
  const int x = is_prologue ? 0 : st->it + 1;
  const unsigned long long word =
      ((unsigned long long)(peer_tag(epoch, x) | flag) << 32) | (unsigned long long)__float_as_uint(acc);
  const size_t off = ((size_t)(x & 1) * world + rank) * xstride + j;
  for (int r = 0; r < world; ++r) {
    const int dst = (rank + r) % world;  // spread the ranks over the links: everyone starts with itself
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peers.buf[dst] + off), "l"(word) : "memory");
  }
}

// reduce_push + finalize of one single-sweep iteration in ONE launch (row-sharded peer loop, default path): thread j
// folds the rank's cluster slabs of column j, stores the tagged word into every peer's buffer, polls the world's
// words of column j in its own buffer, applies the finalize step; the last block (ticket) folds the error partials
// and advances the state machine.  Every block pushes before it polls and the whole grid is co-resident (m / 128
// blocks of 128 threads), so no rank can wait on a word whose producer has not been scheduled.
constexpr int kPeerTailThreads = 128;
__global__ void __launch_bounds__(kPeerTailThreads)
    peer_tail_kernel(State* st, const float* __restrict__ part_sum, int np, size_t stride, int m,
                     const float* __restrict__ b, const float* __restrict__ log2b, float* gs0, float* gs1,
                     double* errpart, float* err_hist, PeerPtrs peers, int world, int rank, size_t xstride,
                     unsigned epoch) {
  if (st->done) return;
  const int j = blockIdx.x * kPeerTailThreads + threadIdx.x;
  const int x = st->it + 1;
  const unsigned tag = peer_tag(epoch, x);
  const int cur = st->cur;
  const float* gcur = cur ? gs1 : gs0;
  float* gnext = cur ? gs0 : gs1;
  const int norm = st->err_norm;
  double e = 0.0;
  int bad = 0;
  if (j < m) {
    float acc = 0.f;
    int p = 0;
    for (; p + 8 <= np; p += 8) {  // eight independent loads in flight, added in slab order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part_sum[(size_t)(p + u) * stride + j];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; p < np; ++p) acc += part_sum[(size_t)p * stride + j];
    if (st->bad) acc = __int_as_float(0x7fc00000);  // a lost sum on this rank stops every rank
    const unsigned long long word = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(acc);
    const size_t par = (size_t)(x & 1) * world;
    for (int r = 0; r < world; ++r) {
      const int dst = (rank + r) % world;  // spread the ranks over the links: everyone starts with itself
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peers.buf[dst] + (par + rank) * xstride + j), "l"(word)
                   : "memory");
    }
    const unsigned long long* mine = peers.buf[rank] + par * xstride + j;
    float sum = 0.f;
    bool any_nan = false;
    for (int r = 0; r < world; ++r) {  // rank order: the same fold on every rank (g bit-equal across ranks)
      const float v = r == rank ? acc : peer_poll(mine + (size_t)r * xstride, tag);
      any_nan |= (v != v);
      sum += v;
    }
    if (any_nan) sum = __int_as_float(0x7fc00000);
    e = finalize_column(sum, log2f(sum), b[j], log2b[j], gcur[j], norm, gnext + j, &bad);
  }
  __shared__ double sh[kPeerTailThreads / 32];
  __shared__ int shbad, is_last;
  if (threadIdx.x == 0) shbad = 0;
  __syncthreads();
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
  if (bad) shbad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < kPeerTailThreads / 32; ++w) tot += sh[w];
    errpart[blockIdx.x] = tot;
    if (shbad) atomicExch(&st->bad, 1);
    __threadfence();
    is_last = (atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double part = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kPeerTailThreads) part += ((volatile double*)errpart)[i];
  part = warp_sum(part);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double tot = 0.0;
  for (int w = 0; w < kPeerTailThreads / 32; ++w) tot += sh[w];
  const float err = (norm == B200OT_NORM_L2) ? (float)sqrt(tot) : (float)tot;
  st->ticket = 0;
  if (((volatile int*)&st->bad)[0]) {  // fast path lost a sum: stop here, the host rewinds and replays robustly
    st->done = 1;
    return;
  }
  State ls = *st;
  advance_state(ls, err, false, err_hist);
  *st = ls;
}

// =============================================================================
// ROBUST kernels (running-max logsumexp, any shape / alignment)
// =============================================================================
struct OnlineLse {
  float mx, s;
  __device__ __forceinline__ void init() {
    mx = -INFINITY;
    s = 0.f;
  }
  __device__ __forceinline__ void add(float x) {
    if (x > mx) {
      s = s * ex2_approx(mx - x) + 1.f;  // mx = -inf: s = 0 * 0 + 1
      mx = x;
    } else if (x > -INFINITY) {
      s += ex2_approx(x - mx);
    }
  }
  __device__ __forceinline__ void merge(float omx, float os) {
    if (omx == -INFINITY) return;
    if (mx == -INFINITY) {
      mx = omx;
      s = os;
      return;
    }
    const float M = fmaxf(mx, omx);
    s = s * ex2_approx(mx - M) + os * ex2_approx(omx - M);
    mx = M;
  }
};

// fs_i = log2 a_i - LSE2_j(gs_j - k C_ij); one CTA per row (grid-stride)
template <bool VEC>
__global__ void __launch_bounds__(256) rowpass_lse_kernel(const float* __restrict__ C, long long ldc,
                                                          int n, int m, State* st, float* fs,
                                                          const float* gs0, const float* gs1,
                                                          const float* __restrict__ a) {
  if (st->done) return;
  const float k = st->kscale;
  const float* __restrict__ gs = st->cur ? gs1 : gs0;
  __shared__ float shm[8], shs[8];
  for (int row = blockIdx.x; row < n; row += gridDim.x) {
    const float* crow = C + (long long)row * ldc;
    OnlineLse o;
    o.init();
    if (VEC) {
      for (int j = threadIdx.x * 4; j < m; j += 256 * 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(crow + j);
        const float4 g4 = *reinterpret_cast<const float4*>(gs + j);
        o.add(fmaf(c4.x, -k, g4.x));
        o.add(fmaf(c4.y, -k, g4.y));
        o.add(fmaf(c4.z, -k, g4.z));
        o.add(fmaf(c4.w, -k, g4.w));
      }
    } else {
      for (int j = threadIdx.x; j < m; j += 256) o.add(fmaf(crow[j], -k, gs[j]));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, o.mx, off);
      const float os = __shfl_xor_sync(0xffffffffu, o.s, off);
      o.merge(omx, os);
    }
    if ((threadIdx.x & 31) == 0) {
      shm[threadIdx.x >> 5] = o.mx;
      shs[threadIdx.x >> 5] = o.s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      OnlineLse t;
      t.init();
      for (int w = 0; w < 8; ++w) t.merge(shm[w], shs[w]);
      const float ai = a[row];
      float fnew = -INFINITY;
      if (ai > 0.f) {
        fnew = log2f(ai) - (t.mx + log2f(t.s));
        if (!(fabsf(fnew) < INFINITY)) atomicExch(&st->bad, 1);
      }
      fs[row] = fnew;
    }
    __syncthreads();
  }
}

// per-column running-max sums of 2^(fs_i + gs_j - k C_ij) over a row range; (max,sum) partial per
// row split.  Threads own columns (4 consecutive in the VEC form) so loads are coalesced.
template <bool VEC>
__global__ void __launch_bounds__(256) colpass_lse_kernel(const float* __restrict__ C, long long ldc,
                                                          int n, int m, const State* st,
                                                          const float* __restrict__ fs,
                                                          const float* gs0, const float* gs1,
                                                          float* __restrict__ part_sum,
                                                          float* __restrict__ part_max, size_t stride,
                                                          int rows_per_split) {
  if (st->done) return;
  const float k = st->kscale;
  const float* __restrict__ gs = st->cur ? gs1 : gs0;
  const int split = blockIdx.y;
  const int r0 = split * rows_per_split;
  int r1 = r0 + rows_per_split;
  r1 = r1 > n ? n : r1;
  if (VEC) {
    const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (j >= m) return;
    const float4 g4 = *reinterpret_cast<const float4*>(gs + j);
    OnlineLse o0, o1, o2, o3;
    o0.init();
    o1.init();
    o2.init();
    o3.init();
    int i = r0;
    for (; i + 4 <= r1; i += 4) {  // four rows in flight per thread (same order of additions as row by row)
      float4 c4[4];
      float fi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        fi[u] = fs[i + u];
        c4[u] = *reinterpret_cast<const float4*>(C + (long long)(i + u) * ldc + j);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o0.add(fmaf(c4[u].x, -k, g4.x + fi[u]));
        o1.add(fmaf(c4[u].y, -k, g4.y + fi[u]));
        o2.add(fmaf(c4[u].z, -k, g4.z + fi[u]));
        o3.add(fmaf(c4[u].w, -k, g4.w + fi[u]));
      }
    }
    for (; i < r1; ++i) {
      const float fi = fs[i];
      const float4 c4 = *reinterpret_cast<const float4*>(C + (long long)i * ldc + j);
      o0.add(fmaf(c4.x, -k, g4.x + fi));
      o1.add(fmaf(c4.y, -k, g4.y + fi));
      o2.add(fmaf(c4.z, -k, g4.z + fi));
      o3.add(fmaf(c4.w, -k, g4.w + fi));
    }
    const size_t o = (size_t)split * stride + j;
    *reinterpret_cast<float4*>(part_sum + o) = make_float4(o0.s, o1.s, o2.s, o3.s);
    *reinterpret_cast<float4*>(part_max + o) = make_float4(o0.mx, o1.mx, o2.mx, o3.mx);
  } else {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= m) return;
    const float gj = gs[j];
    OnlineLse o;
    o.init();
    for (int i = r0; i < r1; ++i) o.add(fmaf(C[(long long)i * ldc + j], -k, gj + fs[i]));
    part_sum[(size_t)split * stride + j] = o.s;
    part_max[(size_t)split * stride + j] = o.mx;
  }
}

// =============================================================================
// snapshot / rewind / export
// =============================================================================
__global__ void snapshot_kernel(State* st, int n, int m, const float* fs, const float* gs0,
                                const float* gs1, float* snap_fs, float* snap_gs) {
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* gs = st->cur ? gs1 : gs0;
  if (i < n) snap_fs[i] = fs[i];
  if (i < m) snap_gs[i] = gs[i];
  if (i == 0) {
    st->snap_it = st->it;
    st->snap_cur = st->cur;
    st->snap_n_err = st->n_err;
    st->snap_err = st->err;
  }
}

__global__ void rewind_kernel(State* st, int n, int m, float* fs, float* gs0, float* gs1,
                              const float* snap_fs, const float* snap_gs) {
  if (st->snap_it < 0) return;  // nothing to go back to (the first g update itself failed): bad / done stay set
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float* gs = st->snap_cur ? gs1 : gs0;
  if (i < n) fs[i] = snap_fs[i];
  if (i < m) gs[i] = snap_gs[i];
  __syncthreads();
  if (i == 0) {
    st->it = st->snap_it;
    st->cur = st->snap_cur;
    st->n_err = st->snap_n_err;
    st->err = st->snap_err;
    st->done = 0;
    st->converged = 0;
    st->bad = 0;
    st->ticket = 0;
    st->stall = 0;
    st->fs_lo[0] = st->fs_lo[1] = INFINITY;  // range unknown after a rewind: per-row shift
    st->fs_hi[0] = st->fs_hi[1] = -INFINITY;
  }
}

__global__ void export_kernel(const State* st, int n, int m, const float* fs, const float* gs0,
                              const float* gs1, float* f, float* g, b200ot_result* res,
                              const float* err_hist, float* err_out, int err_cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float inv = 1.f / st->kscale;
  const float* gs = st->cur ? gs1 : gs0;
  if (f && i < n) f[i] = fs[i] * inv;
  if (g && i < m) g[i] = gs[i] * inv;
  if (err_out && i < err_cap && i < st->n_err && i < kErrHistCap) err_out[i] = err_hist[i];
  if (i == 0 && res) {
    res->n_iter = st->it;
    res->converged = st->converged;
    res->status = st->bad ? B200OT_E_NUMERIC : (st->floor_hit ? 1 : 0);
    res->n_err = st->n_err;
    res->err = st->err;
    res->reserved[0] = res->reserved[1] = res->reserved[2] = 0.f;
  }
}

// =============================================================================
// host side
// =============================================================================
constexpr size_t kSweepSmemMax = 232448 - 1024;

struct FusedCfg {
  int Q, NCH, R, NG, NC, wq;
  size_t smem;
};


template <int NCH, int R>
static cudaError_t fused_set_attr() {
  static PerDeviceOnce attr_once;  // one per instantiation and device
  if (!attr_once.first()) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(sweep_fused_kernel<NCH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(232448 - 1024));
  if (e != cudaSuccess) attr_once.undo();
  return e;
}

template <int NCH, int R>
static cudaError_t launch_fused(const SweepArgs& a, int Q, int NC, size_t smem, cudaStream_t s) {
  cudaError_t e = fused_set_attr<NCH, R>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(NC * Q));
  cfg.blockDim = dim3(kSweepThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)Q;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, sweep_fused_kernel<NCH, R>, a);
}

template <int NCH, int R>
static int query_fused(int Q, size_t smem, int* nc_out) {
  {
    cudaError_t e = fused_set_attr<NCH, R>();
    if (e != cudaSuccess) {
      set_last_cuda_error(e, "cudaFuncSetAttribute(sweep_fused)");
      (void)cudaGetLastError();
      return B200OT_E_LAUNCH;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(Q * 148));
  cfg.blockDim = dim3(kSweepThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)Q;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, sweep_fused_kernel<NCH, R>, &cfg);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "cudaOccupancyMaxActiveClusters");
    return B200OT_E_LAUNCH;
  }
  *nc_out = nc;
  return 0;
}

// rows per group: two register sets of R * 4*NCH exponentials must fit next to the accumulators
static int rows_for_nch(int nch) { return nch <= 2 ? 4 : nch <= 4 ? 2 : 1; }
constexpr int kMaxNch = 6;

// columns per CTA: an even split of m over the cluster, rounded up to 128 floats (512 B) so every CTA's
// slice starts on a cache-line boundary and column validity is warp-uniform; Q = 1 keeps m itself.
static int fused_wq(int m, int Q) { return Q == 1 ? m : ((m + Q - 1) / Q + 127) / 128 * 128; }

static size_t fused_fixed_smem(int R) {
  return 8 * 8 + kXBuf * 8 + (2 * kSweepWarps * R + kXBuf * R * kMaxCluster) * 4 + 128;
}

#define B200OT_DISPATCH_NCH(nch, CALL)            \
  switch (nch) {                                 \
    case 1: CALL(1, 4); break;                   \
    case 2: CALL(2, 4); break;                   \
    case 3: CALL(3, 2); break;                   \
    case 4: CALL(4, 2); break;                   \
    case 5: CALL(5, 1); break;                   \
    default: CALL(6, 1); break;                  \
  }

// Pick the cluster size Q and quads per thread NCH for a width m: among the (Q, NCH) pairs that cover a
// row (Q * NCH * 2048 >= m), take the one that keeps the most SMs busy (the cluster size decides how many
// clusters the GPCs can host: on B200 8 -> 120 SMs, 6 -> 144, 4 -> 132, 1|2 -> 148); ties go to the
// smaller cluster.  B200OT_FUSED_Q / B200OT_FUSED_NCH / B200OT_FUSED_NG override for tuning runs.
static int pick_fused(int n, int m, FusedCfg* out) {
  static int cache_nc[kMaxCluster + 1];  // max co-resident clusters per cluster size (0 = unknown)
  auto clusters_for = [&](int Q, int nch, int* nc) -> int {
    if (cache_nc[Q] > 0) {
      *nc = cache_nc[Q];
      return 0;
    }
    int rc = 0, v = 0;
#define B200OT_Q(NCH_, R_) rc = query_fused<NCH_, R_>(Q, kSweepSmemMax, &v)
    B200OT_DISPATCH_NCH(nch, B200OT_Q)
#undef B200OT_Q
    if (rc) return rc;
    cache_nc[Q] = v > 0 ? v : -1;
    *nc = cache_nc[Q];
    return 0;
  };
  int bestQ = 0, bestN = 0, bestSms = -1;
  const char* eq = getenv("B200OT_FUSED_Q");
  const char* en = getenv("B200OT_FUSED_NCH");
  if (eq && en) {
    const int Q = atoi(eq), N = atoi(en);
    if (Q >= 1 && Q <= kMaxCluster && N >= 1 && N <= kMaxNch && (long long)Q * N * 2048 >= m &&
        N == (fused_wq(m, Q) + 2047) / 2048 && m - (Q - 1) * fused_wq(m, Q) > 2048 * (N - 1)) {
      bestQ = Q;
      bestN = N;
    }
  }
  if (!bestQ) {
    for (int Q = 1; Q <= kMaxCluster; ++Q) {
      const int wq = fused_wq(m, Q);
      const int N = (wq + 2047) / 2048;
      if (N > kMaxNch) continue;
      if (m - (Q - 1) * wq <= 2048 * (N - 1)) continue;  // last CTA must reach into the last quad
      int nc = 0;
      const int rc = clusters_for(Q, N, &nc);
      if (rc) return rc;
      if (nc < 1) continue;
      const int sms = nc * Q;
      if (sms > bestSms) {
        bestSms = sms;
        bestQ = Q;
        bestN = N;
      }
    }
    if (!bestQ) return B200OT_E_UNSUPPORTED;
  }
  const int R = rows_for_nch(bestN);
  const size_t stage = (size_t)R * kSweepThreads * 4 * bestN * 4;
  const size_t fixed = fused_fixed_smem(R);
  int NG = (int)((kSweepSmemMax - fixed) / stage);
  NG = NG > 8 ? 8 : NG;
  if (NG < 2) return B200OT_E_UNSUPPORTED;
  const char* eg = getenv("B200OT_FUSED_NG");
  if (eg && atoi(eg) >= 2 && atoi(eg) <= NG) NG = atoi(eg);
  int NC = 0;
  int rc = clusters_for(bestQ, bestN, &NC);
  if (rc) return rc;
  if (NC < 1) return B200OT_E_UNSUPPORTED;
  const int ngroups = (n + R - 1) / R;
  if (NC > ngroups) NC = ngroups;
  if (NC > kNpCap) NC = kNpCap;
  out->Q = bestQ;
  out->NCH = bestN;
  out->R = R;
  out->NG = NG;
  out->NC = NC;
  out->smem = (size_t)NG * stage + fixed;
  out->wq = fused_wq(m, bestQ);
  return 0;
}

// ---- lite variant host side ---------------------------------------------------------------------
constexpr int kLiteMaxNch = 8;

template <int NCH, bool FUSE>
static cudaError_t lite_set_attr() {
  static PerDeviceOnce attr_once;  // one per instantiation and device
  if (!attr_once.first()) return cudaSuccess;
  cudaError_t e = FUSE ? cudaFuncSetAttribute(sweep_lite_kernel<NCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kLiteSmemMax)
                       : cudaFuncSetAttribute(sweep_lite_plain_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kLiteSmemMax);
  if (e != cudaSuccess) attr_once.undo();
  return e;
}

template <int NCH>
static cudaError_t lite_launch_or_query(const SweepArgs* a, int Q, int NC, size_t smem, cudaStream_t s, int* nc_out) {
  cudaError_t e = lite_set_attr<NCH, false>();
  if (e == cudaSuccess) e = lite_set_attr<NCH, true>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((a ? NC : 2 * 148) * Q));
  cfg.blockDim = dim3(kLiteThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)Q;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (!a) {  // co-resident clusters: the smaller answer of the two instantiations (one grid size serves both)
    int n0 = 0, n1 = 0;
    e = cudaOccupancyMaxActiveClusters(&n0, sweep_lite_plain_kernel<NCH>, &cfg);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveClusters(&n1, sweep_lite_kernel<NCH, true>, &cfg);
    if (e != cudaSuccess) return e;
    *nc_out = n0 < n1 ? n0 : n1;
    return cudaSuccess;
  }
  if (a->fuse) {  // the fused tail holds grid barriers: every CTA must be co-resident, or the launch fails
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, sweep_lite_kernel<NCH, true>, *a);
  }
  SweepArgsPlain pa;
  pa.C = a->C;
  pa.ldc = a->ldc;
  pa.n = a->n;
  pa.m = a->m;
  pa.st = a->st;
  pa.fs = a->fs;
  pa.gs0 = a->gs0;
  pa.gs1 = a->gs1;
  pa.a = a->a;
  pa.part = a->part;
  pa.stride = a->stride;
  pa.ng = a->ng;
  pa.evict_first = a->evict_first;
  pa.wq = a->wq;
  pa.mode = a->mode;
  return cudaLaunchKernelEx(&cfg, sweep_lite_plain_kernel<NCH>, pa);
}

static cudaError_t lite_dispatch(int nch, const SweepArgs* a, int Q, int NC, size_t smem, cudaStream_t s, int* nc_out) {
  switch (nch) {
    case 1: return lite_launch_or_query<1>(a, Q, NC, smem, s, nc_out);
    case 2: return lite_launch_or_query<2>(a, Q, NC, smem, s, nc_out);
    case 3: return lite_launch_or_query<3>(a, Q, NC, smem, s, nc_out);
    case 4: return lite_launch_or_query<4>(a, Q, NC, smem, s, nc_out);
    case 5: return lite_launch_or_query<5>(a, Q, NC, smem, s, nc_out);
    case 6: return lite_launch_or_query<6>(a, Q, NC, smem, s, nc_out);
    case 7: return lite_launch_or_query<7>(a, Q, NC, smem, s, nc_out);
    default: return lite_launch_or_query<8>(a, Q, NC, smem, s, nc_out);
  }
}

// (Q, NCH) for the lite kernel: most co-resident CTAs first, then the smaller cluster
static int pick_lite(int n, int m, FusedCfg* out) {
  static int cache_nc[kMaxCluster + 1][kLiteMaxNch + 1];
  const size_t fixed = 8 * 8 + kXBuf * 8 + (2 * kLiteWarps + kXBuf * kMaxCluster) * 4 + 128;
  int bestQ = 0, bestN = 0, bestCtas = -1, bestNC = 0;
  const char* eq = getenv("B200OT_FUSED_Q");
  const int forceQ = eq ? atoi(eq) : 0;
  for (int Q = 1; Q <= kMaxCluster; ++Q) {
    if (forceQ && Q != forceQ) continue;
    const int wq = fused_wq(m, Q);
    const int N = (wq + 1023) / 1024;
    if (N > kLiteMaxNch) continue;
    if (m - (Q - 1) * wq <= 1024 * (N - 1)) continue;
    if (cache_nc[Q][N] == 0) {
      int v = 0;
      cudaError_t e = lite_dispatch(N, nullptr, Q, 0, kLiteSmemMax, nullptr, &v);
      if (e != cudaSuccess) {
        set_last_cuda_error(e, "lite occupancy query");
        (void)cudaGetLastError();
        return B200OT_E_LAUNCH;
      }
      cache_nc[Q][N] = v > 0 ? v : -1;
    }
    const int nc = cache_nc[Q][N];
    if (nc < 1) continue;
    if (nc * Q > bestCtas) {
      bestCtas = nc * Q;
      bestQ = Q;
      bestN = N;
      bestNC = nc;
    }
  }
  if (!bestQ) return B200OT_E_UNSUPPORTED;
  const size_t stage = (size_t)kLiteThreads * 4 * bestN * 4;
  int NG = (int)((kLiteSmemMax - fixed) / stage);
  NG = NG > 8 ? 8 : NG;
  if (NG < 2) return B200OT_E_UNSUPPORTED;
  const char* eg = getenv("B200OT_FUSED_NG");
  if (eg && atoi(eg) >= 2 && atoi(eg) <= NG) NG = atoi(eg);
  int NC = bestNC;
  if (NC > n) NC = n;
  if (NC > kNpCap) NC = kNpCap;
  out->Q = bestQ;
  out->NCH = bestN;
  out->R = 1;
  out->NG = NG;
  out->NC = NC;
  out->smem = (size_t)NG * stage + fixed;
  out->wq = fused_wq(m, bestQ);
  return 0;
}

static bool fused_eligible(const float* C, int ldc, int n, int m) {
  return (ldc % 4 == 0) && (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && m >= 4 &&
         n >= 1 && (long long)m <= (long long)kMaxCluster * kMaxNch * 2048;
}
static bool vec_eligible(const float* C, int ldc, int m) {
  return (ldc % 4 == 0) && (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
}

// stream C through L2 with an evict-first policy once it cannot stay resident (B200OT_FUSED_EVICT overrides)
static int sweep_evict_first(int n, int m) {
  const char* ev = getenv("B200OT_FUSED_EVICT");
  if (ev) return atoi(ev) ? 1 : 0;
  return ((double)n * (double)m * 4.0 > 100e6) ? 1 : 0;
}

// Peer context of a fused iteration (world = 1: no exchange).  B200OT_FUSE=0 keeps sweep and finalize as separate
// launches (the form the one-GPU emulation of several ranks needs: kernels of different ranks must not wait for
// one another on one device).
struct FuseCtx {
  PeerPtrs peers;
  int world, rank;
  size_t xstride;
  unsigned epoch;
};
static bool g_fuse_broken = false;  // a failed cooperative cluster launch disables the fused form for the process
static long long g_fused_launches = 0, g_fused_fallbacks = 0;  // host-side counters (b200ot_sinkhorn_counter)
static bool fuse_wanted() {
  // Default: separate launches on the plain kernel (measured faster, see sweep_lite_plain_kernel).  B200OT_FUSE=1
  // selects the persistent fused form (one cooperative cluster launch runs sweep, fold, peer exchange, finalize and
  // stopping rule of many iterations).  Read on every call so a test can switch it inside one process.  Nsight
  // Compute cannot launch a cooperative CLUSTER kernel (LaunchFailed under the profiler): never under injection.
  const char* e = getenv("B200OT_FUSE");
  if (!(e && e[0] == '1')) return false;
  if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) return false;
  return !g_fuse_broken;
}

// One sweep (plain form), or -- with a FuseCtx, when the lite kernel applies -- `iters` whole iterations in ONE
// persistent cooperative launch (*fused_out = true).  When the fused form is not available the call launches the
// plain sweep of ONE iteration and the caller finishes it (finalize / exchange) and loops.
static int launch_sweep_fused(const float* C, int ldc, int n, int m, const WsPtrs& w, int* np_out,
                              cudaStream_t s, const FuseCtx* fc = nullptr, bool* fused_out = nullptr, int iters = 1,
                              int* iters_queued = nullptr) {
  if (fused_out) *fused_out = false;
  if (iters_queued) *iters_queued = 1;
  // B200OT_FUSED_VARIANT: "lite" = 256-thread CTAs, two per SM; "pipe" = 512-thread software-pipelined
  static int variant = -1;
  if (variant < 0) {
    const char* ev = getenv("B200OT_FUSED_VARIANT");
    variant = (ev && !strcmp(ev, "pipe")) ? 0 : 1;  // default: lite (measured 1.5x faster at 65536^2)
  }
  const bool lite = variant == 1;
  FusedCfg cfg;
  int rc = lite ? pick_lite(n, m, &cfg) : pick_fused(n, m, &cfg);
  if (rc == B200OT_E_UNSUPPORTED && lite) rc = pick_fused(n, m, &cfg);
  if (rc) return rc;
  SweepArgs a;
  a.C = C;
  a.ldc = ldc;
  a.n = n;
  a.m = m;
  a.st = w.st;
  a.fs = w.fs;
  a.gs0 = w.gs0;
  a.gs1 = w.gs1;
  a.a = w.a;
  a.part = w.part_sum;
  a.stride = w.m_pad;
  a.ng = cfg.NG;
  a.evict_first = sweep_evict_first(n, m);
  a.wq = cfg.wq;
  const char* em = getenv("B200OT_FUSED_MODE");
  a.mode = em ? atoi(em) : 0;
  a.fuse = 0;
  a.iters = 1;
  {
    static int stagger = -1;
    if (stagger < 0) {
      const char* es = getenv("B200OT_STAGGER");
      stagger = es ? atoi(es) : 0;
      if (stagger < 0) stagger = 0;
    }
    a.stagger_ns = stagger;
  }
  cudaError_t e = cudaSuccess;
  if (lite && cfg.R == 1 && cfg.smem <= kLiteSmemMax && cfg.NCH <= kLiteMaxNch && cfg.wq <= 1024 * cfg.NCH) {
    if (fc && fused_out && fuse_wanted() && a.mode == 0) {
      a.fuse = 1;
      a.iters = iters > 0 ? iters : 1;
      {  // the kernel counts the rows of a launch in 32 bits
        const int rows_per_cluster = (n + cfg.NC - 1) / cfg.NC;
        const int cap = (1 << 30) / (rows_per_cluster > 0 ? rows_per_cluster : 1);
        if (a.iters > cap) a.iters = cap > 0 ? cap : 1;
      }
      a.gs0w = w.gs0;
      a.gs1w = w.gs1;
      a.b = w.b;
      a.log2b = w.log2b;
      a.errpart = w.errpart;
      a.err_hist = w.err_hist;
      a.peers = fc->peers;
      a.world = fc->world;
      a.rank = fc->rank;
      a.xstride = fc->xstride;
      a.epoch = fc->epoch;
      e = lite_dispatch(cfg.NCH, &a, cfg.Q, cfg.NC, cfg.smem, s, nullptr);
      if (e == cudaSuccess) {
        *fused_out = true;
        g_fused_launches += a.iters;
        if (iters_queued) *iters_queued = a.iters;
      } else {  // e.g. cooperative + cluster launch refused: fall back to separate launches, for good
        set_last_cuda_error(e, "fused-iteration launch (cooperative cluster launch)");
        (void)cudaGetLastError();
        g_fuse_broken = true;
        ++g_fused_fallbacks;
        a.fuse = 0;
        a.iters = 1;
        e = lite_dispatch(cfg.NCH, &a, cfg.Q, cfg.NC, cfg.smem, s, nullptr);
      }
    } else {
      e = lite_dispatch(cfg.NCH, &a, cfg.Q, cfg.NC, cfg.smem, s, nullptr);
    }
  } else {
#define B200OT_L(NCH_, R_) e = launch_fused<NCH_, R_>(a, cfg.Q, cfg.NC, cfg.smem, s)
    B200OT_DISPATCH_NCH(cfg.NCH, B200OT_L)
#undef B200OT_L
  }
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "sweep_fused launch");
    (void)cudaGetLastError();  // do not leave a stale error for the next launch check
    return B200OT_E_LAUNCH;
  }
  *np_out = cfg.NC;
  return 0;
}

static int launch_rowpass(const float* C, int ldc, int n, int m, const WsPtrs& w, cudaStream_t s) {
  int grid = n < 148 * 8 ? n : 148 * 8;
  if (vec_eligible(C, ldc, m))
    rowpass_lse_kernel<true><<<grid, 256, 0, s>>>(C, ldc, n, m, w.st, w.fs, w.gs0, w.gs1, w.a);
  else
    rowpass_lse_kernel<false><<<grid, 256, 0, s>>>(C, ldc, n, m, w.st, w.fs, w.gs0, w.gs1, w.a);
  B200OT_LAUNCH_OK();
  return 0;
}

static int launch_colpass(const float* C, int ldc, int n, int m, const WsPtrs& w, int* np_out,
                          cudaStream_t s) {
  const bool vec = vec_eligible(C, ldc, m);
  const int cols_per_block = vec ? 1024 : 256;
  const int bx = (m + cols_per_block - 1) / cols_per_block;
  int splits = (148 * 4 + bx - 1) / bx;
  if (splits > kNpCap) splits = kNpCap;
  int rows_per = (n + splits - 1) / splits;
  if (rows_per < 8) rows_per = n < 8 ? n : 8;
  splits = (n + rows_per - 1) / rows_per;
  dim3 grid((unsigned)bx, (unsigned)splits);
  if (vec)
    colpass_lse_kernel<true><<<grid, 256, 0, s>>>(C, ldc, n, m, w.st, w.fs, w.gs0, w.gs1, w.part_sum,
                                                  w.part_max, w.m_pad, rows_per);
  else
    colpass_lse_kernel<false><<<grid, 256, 0, s>>>(C, ldc, n, m, w.st, w.fs, w.gs0, w.gs1, w.part_sum,
                                                   w.part_max, w.m_pad, rows_per);
  B200OT_LAUNCH_OK();
  *np_out = splits;
  return 0;
}

static int launch_finalize(int m, const WsPtrs& w, const float* psum, const float* pmax, int np,
                           size_t stride, int is_prologue, cudaStream_t s,
                           const unsigned long long* tagged = nullptr, unsigned epoch = 0) {
  int groups = np >= 32 ? 8 : np >= 16 ? 4 : np >= 8 ? 2 : 1;
  if (tagged) groups = 1;  // the world's words of a column are gathered by one thread (log-domain combine)
  while (groups > 1 && (long long)(m + kFinalizeThreads / groups - 1) / (kFinalizeThreads / groups) > 4096) groups >>= 1;
  const int cb = kFinalizeThreads / groups;
  const int grid = (m + cb - 1) / cb;
  if (grid > 4096) return B200OT_E_UNSUPPORTED;
  finalize_kernel<<<grid, kFinalizeThreads, 0, s>>>(w.st, psum, pmax, np, stride, m, w.b, w.log2b, w.gs0,
                                                    w.gs1, w.errpart, w.err_hist, is_prologue, groups, tagged, epoch);
  B200OT_LAUNCH_OK();
  return 0;
}

static int resolve_path(int path, const float* C, int ldc, int n, int m) {
  if (path == B200OT_PATH_ROBUST) return B200OT_PATH_ROBUST;
  if (!fused_eligible(C, ldc, n, m)) return B200OT_PATH_ROBUST;
  return B200OT_PATH_FUSED;
}

static int check_problem(const float* C, int ldc, int n, int m, void* ws) {
  if (!C || !ws || n <= 0 || m <= 0 || ldc < m) return B200OT_E_INVALID;
  if ((long long)m > 4096ll * kFinalizeThreads) return B200OT_E_UNSUPPORTED;
  return 0;
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

size_t b200ot_sinkhorn_workspace_bytes(int n, int m) {
  if (n <= 0 || m <= 0) return 0;
  return ws_layout(n, m).total;
}

int b200ot_sinkhorn_setup(int n, int m, const float* a, const float* b, const float* f0,
                          const float* g0, const b200ot_params* prm, void* ws, size_t ws_bytes,
                          void* stream) {
  if (!ws || n <= 0 || m <= 0) return B200OT_E_INVALID;
  if ((long long)m > 4096ll * kFinalizeThreads) return B200OT_E_UNSUPPORTED;
  if (!a || !b || !prm || !(prm->eps > 0.f)) return B200OT_E_INVALID;
  const WsLayout L = ws_layout(n, m);
  if (ws_bytes < L.total) return B200OT_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, L);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int mx = n > m ? n : m;
  init_state_kernel<<<1, 1, 0, s>>>(w.st, *prm);
  B200OT_LAUNCH_OK();
  // the resident kernel's tagged words: zero never matches a tag, so stale workspace bytes cannot be mistaken for data
  if (L.res_ll_bytes) B200OT_CUDA_OK(cudaMemsetAsync(static_cast<char*>(ws) + L.res_ll, 0, L.res_ll_bytes, s));
  init_kernel<<<(mx + 255) / 256, 256, 0, s>>>(w.st, prm->eps, n, m, a, b, f0, g0, w.fs, w.gs0, w.gs1, w.a,
                                               w.b, w.log2b);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_init(const float* C, int ldc, int n, int m, const float* a, const float* b,
                         const float* f0, const float* g0, const b200ot_params* prm, void* ws,
                         size_t ws_bytes, void* stream) {
  int rc = check_problem(C, ldc, n, m, ws);
  if (rc) return rc;
  rc = b200ot_sinkhorn_setup(n, m, a, b, f0, g0, prm, ws, ws_bytes, stream);
  if (rc) return rc;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // first g update from the start potentials (running-max form: safe for any eps)
  int np = 0;
  rc = launch_colpass(C, ldc, n, m, w, &np, s);
  if (rc) return rc;
  return launch_finalize(m, w, w.part_sum, w.part_max, np, w.m_pad, 1, s);
}

int b200ot_sinkhorn_snapshot(int n, int m, void* ws, void* stream) {
  if (!ws || n <= 0 || m <= 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  const int mx = n > m ? n : m;
  snapshot_kernel<<<(mx + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w.st, n, m, w.fs, w.gs0, w.gs1, w.snap_fs, w.snap_gs);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_rewind(int n, int m, void* ws, void* stream) {
  if (!ws || n <= 0 || m <= 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  const int mx = n > m ? n : m;
  // one block per 256 entries; the state reset is done by thread 0 of block 0 after its own
  // copies, the other blocks only copy, and the next kernel on the stream sees everything
  rewind_kernel<<<(mx + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w.st, n, m, w.fs, w.gs0, w.gs1, w.snap_fs, w.snap_gs);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_enqueue(const float* C, int ldc, int n, int m, int iters, int path, void* ws,
                            void* stream) {
  int rc = check_problem(C, ldc, n, m, ws);
  if (rc) return rc;
  if (iters < 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  path = resolve_path(path, C, ldc, n, m);
  rc = b200ot_sinkhorn_snapshot(n, m, ws, stream);
  if (rc) return rc;
  if (path == B200OT_PATH_FUSED && iters > 0) {
    // short iterations: the whole chunk as one persistent launch (resident.cu)
    rc = resident_try_enqueue(C, ldc, n, m, iters, w, s);
    if (rc <= 0) return rc;
  }
  FuseCtx solo;
  memset(&solo, 0, sizeof(solo));
  solo.world = 1;
  for (int i = 0; i < iters; ++i) {
    int np = 0;
    if (path == B200OT_PATH_FUSED) {
      bool fused = false;
      // first try: all remaining iterations as ONE persistent launch (sweeps, folds, finalizes, stopping rule)
      int queued = 1;
      rc = launch_sweep_fused(C, ldc, n, m, w, &np, s, fuse_wanted() ? &solo : nullptr, &fused, iters - i, &queued);
      if (rc) return rc;
      if (fused) {
        i += queued - 1;
        continue;
      }
      rc = launch_finalize(m, w, w.part_sum, nullptr, np, w.m_pad, 0, s);
    } else {
      rc = launch_rowpass(C, ldc, n, m, w, s);
      if (rc) return rc;
      rc = launch_colpass(C, ldc, n, m, w, &np, s);
      if (rc) return rc;
      rc = launch_finalize(m, w, w.part_sum, w.part_max, np, w.m_pad, 0, s);
    }
    if (rc) return rc;
  }
  return 0;
}

int b200ot_sinkhorn_finish(int n, int m, void* ws, float* f, float* g, b200ot_result* result,
                           float* err_hist, int err_hist_cap, void* stream) {
  if (!ws || n <= 0 || m <= 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  int mx = n > m ? n : m;
  if (err_hist && err_hist_cap > mx) mx = err_hist_cap;
  export_kernel<<<(mx + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w.st, n, m, w.fs, w.gs0, w.gs1, f, g, result, w.err_hist, err_hist, err_hist ? err_hist_cap : 0);
  B200OT_LAUNCH_OK();
  return 0;
}

// ---- row panels of one problem (online / C-free solver) -----------------------------------------
// The workspace describes the whole n x m problem; `Cpanel` holds rows [row0, row0 + rows) of the cost, built
// on the fly.  The sweep updates fs[row0 ...] and adds the panel's column sums into s_accum (accumulate = 0
// for the first panel of an iteration).  After the last panel: b200ot_sinkhorn_shard_finalize(n, m, ws, s_accum).
static WsPtrs panel_ptrs(void* ws, int n, int m, int row0) {
  WsPtrs w = ws_ptrs(ws, ws_layout(n, m));
  w.fs += row0;
  w.a += row0;
  return w;
}

int b200ot_sinkhorn_panel_prologue(const float* Cpanel, int ldc, int n, int m, int row0, int rows, void* ws,
                                   float* s_accum, int accumulate, void* stream) {
  int rc = check_problem(Cpanel, ldc, rows, m, ws);
  if (rc) return rc;
  if (!s_accum || row0 < 0 || row0 + rows > n) return B200OT_E_INVALID;
  const WsPtrs w = panel_ptrs(ws, n, m, row0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int np = 0;
  rc = launch_colpass(Cpanel, ldc, rows, m, w, &np, s);
  if (rc) return rc;
  reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, w.part_max, np, w.m_pad, m, s_accum, accumulate);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_panel_sweep(const float* Cpanel, int ldc, int n, int m, int row0, int rows, int path, void* ws,
                                float* s_accum, int accumulate, void* stream) {
  int rc = check_problem(Cpanel, ldc, rows, m, ws);
  if (rc) return rc;
  if (!s_accum || row0 < 0 || row0 + rows > n) return B200OT_E_INVALID;
  const WsPtrs w = panel_ptrs(ws, n, m, row0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  path = resolve_path(path, Cpanel, ldc, rows, m);
  int np = 0;
  if (path == B200OT_PATH_FUSED) {
    rc = launch_sweep_fused(Cpanel, ldc, rows, m, w, &np, s);
    if (rc) return rc;
    reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, nullptr, np, w.m_pad, m, s_accum, accumulate);
  } else {
    rc = launch_rowpass(Cpanel, ldc, rows, m, w, s);
    if (rc) return rc;
    rc = launch_colpass(Cpanel, ldc, rows, m, w, &np, s);
    if (rc) return rc;
    reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, w.part_max, np, w.m_pad, m, s_accum, accumulate);
  }
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_describe(int n, int m, char* buf, int buf_len) {
  if (!buf || buf_len < 8 || n <= 0 || m <= 0) return B200OT_E_INVALID;
  if (m % 4 != 0 || (long long)m > (long long)kMaxCluster * kMaxNch * 2048) {
    snprintf(buf, buf_len, "robust two-sweep kernels (m=%d is not eligible for the single-sweep kernel)", m);
    return 0;
  }
  if (resident_describe(n, m, buf, buf_len)) return 0;
  FusedCfg cfg;
  const char* ev = getenv("B200OT_FUSED_VARIANT");
  bool lite = !(ev && !strcmp(ev, "pipe"));
  int rc = lite ? pick_lite(n, m, &cfg) : pick_fused(n, m, &cfg);
  if (rc == B200OT_E_UNSUPPORTED && lite) {
    lite = false;
    rc = pick_fused(n, m, &cfg);
  }
  if (rc) return rc;
  snprintf(buf, buf_len,
           "%s: cluster=%d CTAs x %d threads, %d clusters (%d CTAs), %d cols/CTA, %d cols/thread, "
           "%d row(s)/group, TMA ring %d x %zu B, smem %zu B/CTA, L2 %s",
           lite ? "sweep_lite_kernel (2 CTAs/SM)" : "sweep_fused_kernel (pipelined, 1 CTA/SM)", cfg.Q,
           lite ? kLiteThreads : kSweepThreads, cfg.NC, cfg.NC * cfg.Q, cfg.wq, 4 * cfg.NCH, cfg.R, cfg.NG,
           (size_t)cfg.R * (lite ? kLiteThreads : kSweepThreads) * 4 * cfg.NCH * 4, cfg.smem,
           sweep_evict_first(n, m) ? "evict-first" : "default policy");
  return 0;
}

long long b200ot_sinkhorn_counter(int which) {
  return which == 0 ? g_fused_launches : which == 1 ? g_fused_fallbacks : -1;
}

int b200ot_sinkhorn_peek(void* ws, int* flags8, void* stream) {
  if (!ws || !flags8) return B200OT_E_INVALID;
  B200OT_CUDA_OK(cudaMemcpyAsync(flags8, ws, 8 * sizeof(int), cudaMemcpyDefault,
                                 static_cast<cudaStream_t>(stream)));
  return 0;
}

int b200ot_sinkhorn_solve(const float* C, int ldc, int n, int m, const float* a, const float* b,
                          const float* f0, const float* g0, const b200ot_params* prm, void* ws,
                          size_t ws_bytes, float* f, float* g, b200ot_result* result_host,
                          float* err_hist, int err_hist_cap, void* stream) {
  if (!prm) return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = b200ot_sinkhorn_init(C, ldc, n, m, a, b, f0, g0, prm, ws, ws_bytes, stream);
  if (rc) return rc;
  // events (and the pinned flag block they guard) belong to the device that was current when they were made:
  // one set per host thread AND device
  static thread_local int* pinned_dev[kMaxDeviceSlots] = {};  // 2 slots x 8 ints + result block
  static thread_local cudaEvent_t ev_dev[kMaxDeviceSlots][2];
  const int slot = device_slot();
  if (!pinned_dev[slot]) {
    B200OT_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&pinned_dev[slot]), 64 * sizeof(int), cudaHostAllocPortable));
    B200OT_CUDA_OK(cudaEventCreateWithFlags(&ev_dev[slot][0], cudaEventDisableTiming));
    B200OT_CUDA_OK(cudaEventCreateWithFlags(&ev_dev[slot][1], cudaEventDisableTiming));
  }
  int* pinned = pinned_dev[slot];
  cudaEvent_t* ev = ev_dev[slot];
  int path = resolve_path(prm->path, C, ldc, n, m);
  const int ce = prm->check_every > 0 ? prm->check_every : 1;
  const int phase = ((prm->check_phase % ce) + ce) % ce;
  // Chunks end on check iterations.  Chunk c is queued before chunk c-1's flags are read, so the
  // host never starves the stream; kernels of a chunk queued after the rule fired are no-ops.
  int it_enq = 0;
  int c = 0;
  bool have_prev = false;
  for (;;) {
    int len = 0;
    if (it_enq < prm->max_iter) {
      len = (((phase - it_enq - 1) % ce) + ce) % ce + 1;
      if (len > prm->max_iter - it_enq) len = prm->max_iter - it_enq;
      // the resident kernel applies the stopping rule itself: one launch runs to convergence or max_iter
      if (path == B200OT_PATH_FUSED && resident_applicable(n, m)) len = prm->max_iter - it_enq;
      rc = b200ot_sinkhorn_enqueue(C, ldc, n, m, len, path, ws, stream);
      if (rc) return rc;
      it_enq += len;
      B200OT_CUDA_OK(cudaMemcpyAsync(pinned + (c & 1) * 8, ws, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
      B200OT_CUDA_OK(cudaEventRecord(ev[c & 1], s));
    }
    if (have_prev) {
      B200OT_CUDA_OK(cudaEventSynchronize(ev[(c - 1) & 1]));
      const int* fl = pinned + ((c - 1) & 1) * 8;
      const int done = fl[1], bad = fl[4];
      if (bad) {
        if (path == B200OT_PATH_ROBUST) break;  // nothing more robust to try; status says so
        // ws holds the snapshot taken at the start of the chunk that lost a sum
        rc = b200ot_sinkhorn_rewind(n, m, ws, stream);
        if (rc) return rc;
        B200OT_CUDA_OK(cudaMemcpyAsync(pinned + 16, ws, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
        B200OT_CUDA_OK(cudaStreamSynchronize(s));
        if (pinned[16 + 4]) break;  // still bad: there was no snapshot to rewind to (status E_NUMERIC)
        it_enq = pinned[16];
        path = B200OT_PATH_ROBUST;
        have_prev = false;
        c = 0;
        continue;
      }
      if (done) break;
    }
    if (len == 0) break;
    have_prev = true;
    ++c;
  }
  b200ot_result* dres = reinterpret_cast<b200ot_result*>(static_cast<char*>(ws) + ws_layout(n, m).errpart);
  // reuse the (now idle) error-partial area as the device-side result block
  rc = b200ot_sinkhorn_finish(n, m, ws, f, g, dres, err_hist, err_hist_cap, stream);
  if (rc) return rc;
  if (result_host) {
    B200OT_CUDA_OK(cudaMemcpyAsync(pinned + 32, dres, sizeof(b200ot_result), cudaMemcpyDeviceToHost, s));
    B200OT_CUDA_OK(cudaStreamSynchronize(s));
    memcpy(result_host, pinned + 32, sizeof(b200ot_result));
  }
  return 0;
}

// ---- row-sharded entry points ------------------------------------------------
int b200ot_sinkhorn_shard_prologue(const float* C, int ldc, int n_local, int m, void* ws,
                                   float* s_local, void* stream) {
  int rc = check_problem(C, ldc, n_local, m, ws);
  if (rc) return rc;
  if (!s_local) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int np = 0;
  rc = launch_colpass(C, ldc, n_local, m, w, &np, s);
  if (rc) return rc;
  reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, w.part_max, np, w.m_pad, m, s_local, 0);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_shard_sweep(const float* C, int ldc, int n_local, int m, int path, void* ws,
                                float* s_local, void* stream) {
  int rc = check_problem(C, ldc, n_local, m, ws);
  if (rc) return rc;
  if (!s_local) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  path = resolve_path(path, C, ldc, n_local, m);
  int np = 0;
  if (path == B200OT_PATH_FUSED) {
    rc = launch_sweep_fused(C, ldc, n_local, m, w, &np, s);
    if (rc) return rc;
    reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, nullptr, np, w.m_pad, m, s_local, 0);
  } else {
    rc = launch_rowpass(C, ldc, n_local, m, w, s);
    if (rc) return rc;
    rc = launch_colpass(C, ldc, n_local, m, w, &np, s);
    if (rc) return rc;
    reduce_parts_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, w.part_max, np, w.m_pad, m, s_local, 0);
  }
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_shard_finalize(int n_local, int m, void* ws, const float* s_total,
                                   int is_prologue, void* stream) {
  if (!ws || !s_total || n_local <= 0 || m <= 0) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  return launch_finalize(m, w, s_total, nullptr, 1, 0, is_prologue, static_cast<cudaStream_t>(stream));
}

// ---- row-sharded entry points, peer-memory form (no NCCL in the loop) ---------------------------------------
size_t b200ot_peer_exchange_bytes(int world, int m) {
  if (world < 1 || world > kMaxPeers || m <= 0) return 0;
  return (size_t)2 * world * align_up((size_t)m, 64) * sizeof(unsigned long long);
}

int b200ot_sinkhorn_shard_push(const float* C, int ldc, int n_local, int m, int path, void* ws,
                               void* const* peer_bufs, int world, int rank, unsigned epoch, int is_prologue,
                               void* stream) {
  int rc = check_problem(C, ldc, n_local, m, ws);
  if (rc) return rc;
  if (!peer_bufs || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PeerPtrs pp;
  for (int r = 0; r < kMaxPeers; ++r) pp.buf[r] = r < world ? static_cast<unsigned long long*>(peer_bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pp.buf[r]) return B200OT_E_INVALID;
  int np = 0;
  const float* pmax = nullptr;
  if (is_prologue) {
    rc = launch_colpass(C, ldc, n_local, m, w, &np, s);
    pmax = w.part_max;
  } else {
    path = resolve_path(path, C, ldc, n_local, m);
    if (path == B200OT_PATH_FUSED) {
      rc = launch_sweep_fused(C, ldc, n_local, m, w, &np, s);
    } else {
      rc = launch_rowpass(C, ldc, n_local, m, w, s);
      if (rc) return rc;
      rc = launch_colpass(C, ldc, n_local, m, w, &np, s);
      pmax = w.part_max;
    }
  }
  if (rc) return rc;
  reduce_push_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, pmax, np, w.m_pad, m, pp, world, rank,
                                                     align_up((size_t)m, 64), epoch & 0x7ffu, is_prologue);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_sinkhorn_shard_finalize_peer(int n_local, int m, void* ws, const void* my_buf, int world, unsigned epoch,
                                        int is_prologue, void* stream) {
  if (!ws || !my_buf || n_local <= 0 || m <= 0 || world < 1 || world > kMaxPeers) return B200OT_E_INVALID;
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  return launch_finalize(m, w, nullptr, nullptr, world, align_up((size_t)m, 64), is_prologue,
                         static_cast<cudaStream_t>(stream), static_cast<const unsigned long long*>(my_buf),
                         epoch & 0x7ffu);
}

// B200OT_PEER_TAIL=0: keep reduce_push and finalize as two launches (A/B switch; re-read on every call)
static bool peer_tail_wanted() {
  const char* e = getenv("B200OT_PEER_TAIL");
  return !(e && e[0] == '0');
}

int b200ot_sinkhorn_shard_run_peer(const float* C, int ldc, int n_local, int m, int iters, int path, void* ws,
                                   void* const* peer_bufs, int world, int rank, unsigned epoch, void* stream) {
  if (iters < 0 || !peer_bufs || rank < 0 || rank >= world || world > kMaxPeers) return B200OT_E_INVALID;
  int rc = check_problem(C, ldc, n_local, m, ws);
  if (rc) return rc;
  const bool single_sweep = resolve_path(path, C, ldc, n_local, m) == B200OT_PATH_FUSED;
  const bool can_fuse = single_sweep && fuse_wanted();
  const WsPtrs w = ws_ptrs(ws, ws_layout(n_local, m));
  FuseCtx fc;
  memset(&fc, 0, sizeof(fc));
  for (int r = 0; r < world; ++r) {
    if (!peer_bufs[r]) return B200OT_E_INVALID;
    fc.peers.buf[r] = static_cast<unsigned long long*>(peer_bufs[r]);
  }
  fc.world = world;
  fc.rank = rank;
  fc.xstride = align_up((size_t)m, 64);
  fc.epoch = epoch & 0x7ffu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < iters; ++i) {
    bool fused = false;
    if (can_fuse && !g_fuse_broken) {
      // ONE persistent launch for all `iters` iterations: sweep + fold + push to the peers + poll + finalize + state
      // machine, the TMA ring running across the iteration boundaries
      int np = 0;
      int queued = 1;
      rc = launch_sweep_fused(C, ldc, n_local, m, w, &np, s, &fc, &fused, iters - i, &queued);
      if (rc) return rc;
      if (fused) {
        i += queued - 1;
        continue;
      }
      {  // the plain sweep of one iteration ran instead (no fused tail): finish it the unfused way
        PeerPtrs pp = fc.peers;
        reduce_push_kernel<<<(m + 255) / 256, 256, 0, s>>>(w.st, w.part_sum, nullptr, np, w.m_pad, m, pp, world, rank,
                                                           fc.xstride, fc.epoch, 0);
        B200OT_LAUNCH_OK();
        rc = b200ot_sinkhorn_shard_finalize_peer(n_local, m, ws, peer_bufs[rank], world, epoch, 0, stream);
        if (rc) return rc;
      }
      continue;
    }
    if (single_sweep && m <= 4096 * kPeerTailThreads && peer_tail_wanted()) {
      // default: the plain sweep, then ONE launch that folds, pushes, polls and finalizes
      int np = 0;
      rc = launch_sweep_fused(C, ldc, n_local, m, w, &np, s);
      if (rc) return rc;
      peer_tail_kernel<<<(m + kPeerTailThreads - 1) / kPeerTailThreads, kPeerTailThreads, 0, s>>>(
          w.st, w.part_sum, np, w.m_pad, m, w.b, w.log2b, w.gs0, w.gs1, w.errpart, w.err_hist, fc.peers, world, rank,
          fc.xstride, fc.epoch);
      B200OT_LAUNCH_OK();
      continue;
    }
    rc = b200ot_sinkhorn_shard_push(C, ldc, n_local, m, path, ws, peer_bufs, world, rank, epoch, 0, stream);
    if (rc) return rc;
    rc = b200ot_sinkhorn_shard_finalize_peer(n_local, m, ws, peer_bufs[rank], world, epoch, 0, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
