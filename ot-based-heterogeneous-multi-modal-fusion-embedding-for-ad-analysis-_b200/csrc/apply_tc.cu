// Plan application on the 5th-generation tensor cores (tcgen05 + TMEM), single pass over C.
//
//   Z = P V   (rows of Z = rows i of C)       P_ij = exp((f_i + g_j - C_ij)/eps)
//   Z = P^T U (rows of Z = columns j of C)
//
// is the barycentric projection `(T / rowsum) @ Y` of perturbot/perturbot/eval/match.py:202-206, the
// reference's `pet @ T.t()` (MRI_PET_OT_OT_per_epoch_attn.py:728) and both halves of the envelope gradient
// dX = 2 (diag(P1) X - P Y), dY = 2 (diag(P^T 1) Y - P^T X).  P is never materialised: every element of C is read
// ONCE, its plan entry is evaluated once (one FFMA + one MUFU.EX2) by the producer warps, split into two bf16
// parts p = p1 + p2 and written to shared memory as the A operand of tcgen05.mma; the right-hand side V is split
// the same way by a pre-pass (v = v1 + v2) and pre-tiled so that TMA bulk copies drop it into shared memory as the
// B operand.  The product is the 3-term split p1 v2 + p2 v1 + p1 v1 (error ~2^-16 per product, far inside the
// 1e-4 the north star asks of the fused embedding) accumulated in fp32 in TMEM: one CTA owns a 128-row block of Z
// with all dv <= 512 accumulator columns resident (128 lanes x 512 columns = the whole TMEM of the SM), so C is
// streamed exactly once whatever dv is.  Row sums of P (for the normalised projection and the envelope gradient)
// are accumulated in fp32 by the producers from the same registers.
//
// CTA = 10 warps, persistent over (row block, K split) work items:
//   warp 0      TMA producer (one lane): ring of 2 stages of the pre-tiled V parts (<= 64 KiB per stage)
//   warp 1      MMA issuer   (one lane): tcgen05.mma cta_group::1 kind::f16, M = 128, N <= 256 (x2 halves), K = 16
//   warps 2-9   plan producers: C (coalesced global loads, software-prefetched one stage ahead) -> 2^x -> bf16
//               split -> st.shared (canonical no-swizzle core matrices: K-major for P V, MN-major for P^T U, so
//               both forms read C along its rows) -> fence.proxy.async -> mbarrier;
//               after the last K step of an item the same warps are the epilogue (tcgen05.ld -> scale -> stores).
// Small outputs (few row blocks) split the K range over several CTAs; the partial accumulators go to the
// workspace and a small second kernel folds them in fixed order (bit-reproducible, no atomics).
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200ot {

constexpr int AP_BM = 128;      // rows of Z per CTA tile (UMMA M)
constexpr int AP_KC = 32;       // plan columns (K) per stage
constexpr int AP_NH = 256;      // max UMMA N; dv <= 2 * AP_NH
constexpr int AP_STAGES = 2;
constexpr int AP_PARTS = 2;     // bf16 parts of P and of V
constexpr int AP_PROD_WARPS = 8;
constexpr int AP_THREADS = 64 + AP_PROD_WARPS * 32;
constexpr uint32_t AP_A_TILE = AP_BM * AP_KC * 2;  // one part of the plan tile: 8 KiB
constexpr int AP_MAX_DV = 2 * AP_NH;

struct ApplySub {
  const float* C;       // n x m cost, row-major
  long long ldc;
  const float* frow;    // potential of the OUTPUT rows (f for P V, g for P^T U), natural units
  const float* fk;      // potential of the contracted index (g for P V, f for P^T U)
  const uint8_t* Vt;    // pre-tiled bf16 parts of the right-hand side
  const float* W;       // mode 2: Z = scale * (rowsum * W - acc)
  long long ldw;
  float* Z;
  long long ldz;
  float* rowsum_out;    // optional: rows floats
  float* part;          // split-K partial accumulators [S][rows_pad][ldp]
  float* part_rs;       // [S][rows_pad]
  int rows, kdim;       // output rows, contracted length
  int transposed;       // 0: output rows are rows of C; 1: output rows are columns of C
  int nkb;              // ceil(kdim / AP_KC)
  int dv, dv16, TR, nhalves, ldp;
  int S, kb_per_split, nblocks;
  int mode;             // 0 raw, 1 divide by the row sum (0 -> 1e-30), 2 envelope form
  float scale;
  // weighted plan (implicit differentiation, C-weighted products): the contracted entries are
  // W = P * (w0 + w1 * C + wrow[output row] + wk[contracted index]) instead of P; row sums are those of W
  int weighted;
  float w0, w1;
  const float* wrow;    // per output row (may be null)
  const float* wk;      // per contracted index (may be null)
  int vec;              // 16-byte loads along the rows of C are legal
};

struct ApplyArgs {
  ApplySub sub[2];
  int nsub;
  int items[2];  // work items (row blocks x splits) of each sub-problem
  float k;       // log2(e) / eps
};

// ---- pre-pass: right-hand side (rows_in x dv, fp32) -> two bf16 parts, pre-tiled ----------------------------
// Tile (kb, part, half) holds TR "rows" (columns c of V, i.e. the N index of the MMA) x 32 K values (rows of V),
// K-major core matrices: 16-byte unit (c, kc) at ((kc * TR/8) + c/8) * 128 + (c % 8) * 16 holds K = kc*8 .. +7.
__global__ void __launch_bounds__(256) apply_split_v_kernel(const float* __restrict__ V, long long ldv, int rows_in,
                                                            int dv, int TR, int nhalves, int nkb,
                                                            uint8_t* __restrict__ out) {
  const int ncols = nhalves * TR;
  const long long units = (long long)nkb * 4 * ncols;
  const size_t tile_bytes = (size_t)TR * AP_KC * 2;
  for (long long u = (long long)blockIdx.x * 256 + threadIdx.x; u < units; u += (long long)gridDim.x * 256) {
    const int cc = (int)(u % ncols);
    const long long t = u / ncols;
    const int kc = (int)(t & 3);
    const long long kb = t >> 2;
    const int half = cc / TR, c = cc - half * TR;
    __nv_bfloat16 p1[8], p2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long r = kb * AP_KC + kc * 8 + e;
      const float v = (r < rows_in && cc < dv) ? V[r * ldv + cc] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      p1[e] = h;
      p2[e] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    const size_t inner = ((size_t)kc * (TR / 8) + (c >> 3)) * 128 + (size_t)(c & 7) * 16;
    const size_t tile0 = ((size_t)kb * AP_PARTS + 0) * nhalves + half;
    const size_t tile1 = ((size_t)kb * AP_PARTS + 1) * nhalves + half;
    *reinterpret_cast<uint4*>(out + tile0 * tile_bytes + inner) = *reinterpret_cast<const uint4*>(p1);
    *reinterpret_cast<uint4*>(out + tile1 * tile_bytes + inner) = *reinterpret_cast<const uint4*>(p2);
  }
}

// ---- helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// p = p1 + p2 for a pair: returns the packed first parts, writes the packed second parts
__device__ __forceinline__ uint32_t split_pair(float a, float b, uint32_t& second) {
  const uint32_t h = pack_bf16x2(a, b);
  const float ra = a - __uint_as_float(h << 16);
  const float rb = b - __uint_as_float(h & 0xffff0000u);
  second = pack_bf16x2(ra, rb);
  return h;
}
__device__ __forceinline__ void bar_sync_producers() {
  asm volatile("bar.sync 1, %0;" ::"n"(AP_PROD_WARPS * 32) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- the kernel ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AP_THREADS, 1) apply_tc_kernel(const __grid_constant__ ApplyArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: A[stage][part] (8 KiB each) | V[stage] (vstage bytes each) | row sums | barriers
  const ApplySub& s0 = p.sub[0];
  uint32_t vstage = (uint32_t)(AP_PARTS * s0.nhalves * s0.TR * AP_KC * 2);
  if (p.nsub > 1) {
    const uint32_t v1 = (uint32_t)(AP_PARTS * p.sub[1].nhalves * p.sub[1].TR * AP_KC * 2);
    vstage = v1 > vstage ? v1 : vstage;
  }
  uint8_t* sA = smem;
  uint8_t* sV = smem + AP_STAGES * AP_PARTS * AP_A_TILE;
  float* s_rs = reinterpret_cast<float*>(sV + (size_t)AP_STAGES * vstage);  // [2 parities][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rs + 2 * 2 * AP_BM);
  uint64_t* vfull = bars;                      // [AP_STAGES] TMA -> MMA
  uint64_t* pfull = bars + AP_STAGES;          // [AP_STAGES] producers -> MMA
  uint64_t* empty = bars + 2 * AP_STAGES;      // [AP_STAGES] MMA -> TMA, producers
  uint64_t* accfull = bars + 3 * AP_STAGES;    // MMA -> epilogue
  uint64_t* accempty = accfull + 1;            // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = p.items[0] + (p.nsub > 1 ? p.items[1] : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < AP_STAGES; ++s) {
      mbar_init(smem_u32(vfull + s), 1);
      mbar_init(smem_u32(pfull + s), AP_PROD_WARPS);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(accfull), 1);
    mbar_init(smem_u32(accempty), AP_PROD_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (sub-problem, row block, K split)
  auto decode = [&](int item, int& si, int& rb, int& kb0, int& kb1, int& split) {
    si = (item >= p.items[0]) ? 1 : 0;
    const ApplySub& sp = p.sub[si];
    const int loc = item - (si ? p.items[0] : 0);
    rb = loc / sp.S;
    split = loc - rb * sp.S;
    kb0 = split * sp.kb_per_split;
    kb1 = kb0 + sp.kb_per_split;
    kb1 = kb1 > sp.nkb ? sp.nkb : kb1;
  };

  if (warp == 0) {
    // ===================== TMA producer: V tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int si, rb, kb0, kb1, split;
        decode(item, si, rb, kb0, kb1, split);
        const ApplySub& sp = p.sub[si];
        const uint32_t bytes = (uint32_t)(AP_PARTS * sp.nhalves * sp.TR * AP_KC * 2);
        const uint32_t tile = (uint32_t)(sp.TR * AP_KC * 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(empty + stage), phase ^ 1);
          const uint32_t bar = smem_u32(vfull + stage);
          mbar_arrive_expect_tx(bar, bytes);
          const uint8_t* src = sp.Vt + (size_t)kb * bytes;
          const uint32_t dst = smem_u32(sV + (size_t)stage * vstage);
          for (uint32_t t = 0; t < (uint32_t)(AP_PARTS * sp.nhalves); ++t)
            bulk_g2s(dst + t * tile, src + (size_t)t * tile, tile, bar);
          if (++stage == AP_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int si, rb, kb0, kb1, split;
        decode(item, si, rb, kb0, kb1, split);
        const ApplySub& sp = p.sub[si];
        const uint32_t TR = (uint32_t)sp.TR;
        const uint32_t tile = TR * AP_KC * 2;
        const int n0 = sp.dv16 > AP_NH ? AP_NH : sp.dv16;
        const int n1 = sp.dv16 - n0;
        // instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N per half
        // forward: A is K-major (unit (row r, K chunk u) at (u * 16 + r / 8) * 128 + (r % 8) * 16: LBO = 2048 between
        // the two K chunks of one MMA, SBO = 128 between 8-row groups, 4096 per K = 16 step); transposed: A is
        // MN-major (idesc bit 15; unit (jg, i) at jg * 512 + (i / 8) * 128 + (i % 8) * 16: LBO = 128 between K
        // groups, SBO = 512 between groups of 8 output rows, 256 per K = 16 step)
        const uint32_t a_lbo = sp.transposed ? 128u : AP_BM * 16u;
        const uint32_t a_sbo = sp.transposed ? 512u : 128u;
        const uint32_t a_kstep = sp.transposed ? 256u : 2u * AP_BM * 16u;
        const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AP_BM >> 4) << 24) |
                                    (sp.transposed ? (1u << 15) : 0u);
        const uint32_t idesc0 = idesc_base | ((uint32_t)(n0 >> 3) << 17);
        const uint32_t idesc1 = idesc_base | ((uint32_t)(n1 >> 3) << 17);
        mbar_wait(smem_u32(accempty), aphase ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(vfull + stage), phase);
          mbar_wait(smem_u32(pfull + stage), phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)stage * AP_PARTS * AP_A_TILE);
          const uint32_t v0 = smem_u32(sV + (size_t)stage * vstage);
#pragma unroll
          for (int k16 = 0; k16 < AP_KC / 16; ++k16) {
            for (int h = 0; h < sp.nhalves; ++h) {
              const uint32_t d = tmem_base + (uint32_t)h * AP_NH;
              const uint32_t idesc = h ? idesc1 : idesc0;
              const bool first = (kb == kb0 && k16 == 0);
              // terms: p1.v2, p2.v1, p1.v1  (A part, V part)
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                const uint32_t pa = (t == 1) ? 1u : 0u;
                const uint32_t pb = (t == 0) ? 1u : 0u;
                const uint64_t ad = make_smem_desc(a0 + pa * AP_A_TILE + (uint32_t)k16 * a_kstep, a_lbo, a_sbo);
                const uint64_t bd = make_smem_desc(v0 + (pb * (uint32_t)sp.nhalves + (uint32_t)h) * tile +
                                                       (uint32_t)k16 * 2u * (TR * 16u), TR * 16u, 128u);
                tc_mma_bf16(d, ad, bd, idesc, (first && t == 0) ? 0u : 1u);
              }
            }
          }
          tc_commit(smem_u32(empty + stage));  // the stage is free once these MMAs have read it
          if (++stage == AP_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(accfull));
        aphase ^= 1;
      }
    }
  } else {
    // ===================== plan producers + epilogue (warps 2..9) =====================
    const int pw = warp - 2;
    const float k = p.k;
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    int parity = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, parity ^= 1) {
      int si, rb, kb0, kb1, split;
      decode(item, si, rb, kb0, kb1, split);
      const ApplySub& sp = p.sub[si];
      float* rs_buf = s_rs + parity * (2 * AP_BM);
      if (!sp.transposed) {
        // thread = (tile rows r0 and r0 + 64, K quarter kc): columns kc*4..+3 and 16 + kc*4..+3 of the stage
        // lane = (8-byte half of the unit, row % 8, unit): a half-warp stores 128 contiguous bytes
        const int r0 = pw * 8 + ((lane >> 1) & 7);
        const int kc = (lane & 1) | ((lane >> 4) << 1);
        const long long rowA = (long long)rb * AP_BM + r0, rowB = rowA + 64;
        const bool okA = rowA < sp.rows, okB = rowB < sp.rows;
        const float fA = okA ? sp.frow[rowA] * k : 0.f;
        const float fB = okB ? sp.frow[rowB] * k : 0.f;
        const float wA = (sp.weighted && sp.wrow && okA) ? sp.wrow[rowA] : 0.f;
        const float wB = (sp.weighted && sp.wrow && okB) ? sp.wrow[rowB] : 0.f;
        const float* cA = sp.C + (okA ? rowA : (long long)sp.rows - 1) * sp.ldc;
        const float* cB = sp.C + (okB ? rowB : (long long)sp.rows - 1) * sp.ldc;
        float rsA = 0.f, rsB = 0.f;
        float nA[8], nB[8], nG[8];
        auto load = [&](int kb) {
          const int j0 = kb * AP_KC + kc * 4;
          if (sp.vec && j0 + 20 <= sp.kdim) {
            const float4 a0 = __ldcs(reinterpret_cast<const float4*>(cA + j0));
            const float4 a1 = __ldcs(reinterpret_cast<const float4*>(cA + j0 + 16));
            const float4 b0 = __ldcs(reinterpret_cast<const float4*>(cB + j0));
            const float4 b1 = __ldcs(reinterpret_cast<const float4*>(cB + j0 + 16));
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(sp.fk + j0));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(sp.fk + j0 + 16));
            nA[0] = a0.x, nA[1] = a0.y, nA[2] = a0.z, nA[3] = a0.w, nA[4] = a1.x, nA[5] = a1.y, nA[6] = a1.z, nA[7] = a1.w;
            nB[0] = b0.x, nB[1] = b0.y, nB[2] = b0.z, nB[3] = b0.w, nB[4] = b1.x, nB[5] = b1.y, nB[6] = b1.z, nB[7] = b1.w;
            nG[0] = g0.x, nG[1] = g0.y, nG[2] = g0.z, nG[3] = g0.w, nG[4] = g1.x, nG[5] = g1.y, nG[6] = g1.z, nG[7] = g1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = j0 + (e & 3) + (e >> 2) * 16;
              const bool ok = j < sp.kdim;
              nA[e] = ok ? cA[j] : 0.f;
              nB[e] = ok ? cB[j] : 0.f;
              nG[e] = ok ? sp.fk[j] : -INFINITY;  // no mass beyond the last column
            }
          }
        };
        if (kb0 < kb1) load(kb0);
        for (int kb = kb0; kb < kb1; ++kb) {
          float cAv[8], cBv[8], gv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) cAv[e] = nA[e], cBv[e] = nB[e], gv[e] = nG[e];
          if (kb + 1 < kb1) load(kb + 1);
          float pa[8], pb[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float gs = gv[e] * k;
            pa[e] = okA ? ex2_approx(fmaf(cAv[e], -k, gs + fA)) : 0.f;
            pb[e] = okB ? ex2_approx(fmaf(cBv[e], -k, gs + fB)) : 0.f;
          }
          if (sp.weighted) {  // CTA-uniform
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = kb * AP_KC + kc * 4 + (e & 3) + (e >> 2) * 16;
              const float wj = (sp.wk && j < sp.kdim) ? __ldg(sp.wk + j) : 0.f;
              pa[e] *= fmaf(sp.w1, cAv[e], sp.w0 + wA + wj);
              pb[e] *= fmaf(sp.w1, cBv[e], sp.w0 + wB + wj);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            rsA += pa[e];
            rsB += pb[e];
          }
          mbar_wait(smem_u32(empty + stage), phase ^ 1);
          const uint32_t a0 = smem_u32(sA + (size_t)stage * AP_PARTS * AP_A_TILE);
          // 16-byte unit (row r, K chunk u) at (u * 16 + r / 8) * 128 + (r % 8) * 16; this thread owns the
          // 8-byte half (kc & 1) of units (kc >> 1) and 2 + (kc >> 1)
          const uint32_t offA = (uint32_t)(((kc >> 1) * 16 + (r0 >> 3)) * 128 + (r0 & 7) * 16 + (kc & 1) * 8);
          const uint32_t offB = offA + 8 * 128;  // row + 64 = 8 row groups further
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t s0, s1, t0, t1;
            const uint32_t h0 = split_pair(pa[hlf * 4 + 0], pa[hlf * 4 + 1], s0);
            const uint32_t h1 = split_pair(pa[hlf * 4 + 2], pa[hlf * 4 + 3], s1);
            const uint32_t g0 = split_pair(pb[hlf * 4 + 0], pb[hlf * 4 + 1], t0);
            const uint32_t g1 = split_pair(pb[hlf * 4 + 2], pb[hlf * 4 + 3], t1);
            const uint32_t uo = (uint32_t)hlf * (2 * 16 * 128);  // units +2
            st_shared_v2(a0 + offA + uo, h0, h1);
            st_shared_v2(a0 + AP_A_TILE + offA + uo, s0, s1);
            st_shared_v2(a0 + offB + uo, g0, g1);
            st_shared_v2(a0 + AP_A_TILE + offB + uo, t0, t1);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(pfull + stage));
          if (++stage == AP_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        rsA += __shfl_xor_sync(0xffffffffu, rsA, 1);  // the four K quarters of a row sit in lane bits 0 and 4
        rsA += __shfl_xor_sync(0xffffffffu, rsA, 16);
        rsB += __shfl_xor_sync(0xffffffffu, rsB, 1);
        rsB += __shfl_xor_sync(0xffffffffu, rsB, 16);
        if (kc == 0) {
          rs_buf[r0] = rsA;
          rs_buf[r0 + 64] = rsB;
          rs_buf[AP_BM + r0] = 0.f;
          rs_buf[AP_BM + r0 + 64] = 0.f;
        }
      } else {
        // Transposed (Z = P^T U): output rows are columns j of C, the contraction runs over rows i.  The plan
        // tile is written MN-major (8 consecutive j of one row i form a 16-byte unit; unit (jg, i) at
        // jg * 512 + (i / 8) * 128 + (i % 8) * 16), so the loads stay coalesced along the rows of C exactly as
        // in the forward form.
        // lane = (8-byte half of the unit, i % 8, low bit of the K group): consecutive lanes store consecutive
        // addresses (ncu: the (i % 8, half, K group) order cost 7.9 M shared-store bank conflicts per launch at
        // 16384^2, this order 1.4 M in the forward form)
        const int half = lane & 1, i8 = (lane >> 1) & 7, kgbit = lane >> 4;
        const long long jb = (long long)rb * AP_BM;
        float gJ[2][4], rs[2][4], wJ[2][4];
        bool vecok[2];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const long long j0 = jb + (pw * 2 + jj) * 8 + half * 4;
          vecok[jj] = sp.vec && (j0 + 4 <= sp.rows);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            gJ[jj][e] = (j0 + e < sp.rows) ? sp.frow[j0 + e] * k : -INFINITY;  // no mass beyond the last column
            wJ[jj][e] = (sp.weighted && sp.wrow && j0 + e < sp.rows) ? sp.wrow[j0 + e] : 0.f;
            rs[jj][e] = 0.f;
          }
        }
        float nC[4][4], nF[2];
        auto load = [&](int kb) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const long long i = (long long)kb * AP_KC + (kgbit + 2 * kk) * 8 + i8;
            nF[kk] = i < sp.kdim ? __ldg(sp.fk + i) : -INFINITY;  // scaled at the point of use: no wait inside the prefetch
            const float* crow = sp.C + (i < sp.kdim ? i : (long long)sp.kdim - 1) * sp.ldc;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const long long j0 = jb + (pw * 2 + jj) * 8 + half * 4;
              if (vecok[jj]) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(crow + j0));
                nC[jj * 2 + kk][0] = v.x, nC[jj * 2 + kk][1] = v.y, nC[jj * 2 + kk][2] = v.z, nC[jj * 2 + kk][3] = v.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) nC[jj * 2 + kk][e] = (j0 + e < sp.rows) ? crow[j0 + e] : 0.f;
              }
            }
          }
        };
        if (kb0 < kb1) load(kb0);
        for (int kb = kb0; kb < kb1; ++kb) {
          float cv[4][4], fv[2];
#pragma unroll
          for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) cv[t][e] = nC[t][e];
          fv[0] = nF[0] * k, fv[1] = nF[1] * k;
          if (kb + 1 < kb1) load(kb + 1);
          float pv[4][4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float v = ex2_approx(fmaf(cv[jj * 2 + kk][e], -k, fv[kk] + gJ[jj][e]));
                if (sp.weighted) {  // CTA-uniform
                  const long long i = (long long)kb * AP_KC + (kgbit + 2 * kk) * 8 + i8;
                  const float wi = (sp.wk && i < sp.kdim) ? __ldg(sp.wk + i) : 0.f;
                  v *= fmaf(sp.w1, cv[jj * 2 + kk][e], sp.w0 + wJ[jj][e] + wi);
                }
                pv[jj * 2 + kk][e] = v;
                rs[jj][e] += v;
              }
          mbar_wait(smem_u32(empty + stage), phase ^ 1);
          const uint32_t a0 = smem_u32(sA + (size_t)stage * AP_PARTS * AP_A_TILE);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t off = (uint32_t)((pw * 2 + jj) * 512 + (kgbit + 2 * kk) * 128 + i8 * 16 + half * 8);
              uint32_t s0, s1;
              const uint32_t h0 = split_pair(pv[jj * 2 + kk][0], pv[jj * 2 + kk][1], s0);
              const uint32_t h1 = split_pair(pv[jj * 2 + kk][2], pv[jj * 2 + kk][3], s1);
              st_shared_v2(a0 + off, h0, h1);
              st_shared_v2(a0 + AP_A_TILE + off, s0, s1);
            }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(pfull + stage));
          if (++stage == AP_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        // column sums: fold over the lanes that share (unit half) -- bits 1..3 (i % 8) and 4 (K group)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = rs[jj][e];
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (i8 == 0 && kgbit == 0) {
              const int r = (pw * 2 + jj) * 8 + half * 4 + e;
              rs_buf[r] = v;
              rs_buf[AP_BM + r] = 0.f;
            }
          }
      }
      bar_sync_producers();  // row sums complete

      // ---- epilogue: TMEM -> registers -> global ----
      const int q = warp & 3;    // TMEM lane quarter this warp may read
      const int grp = pw >> 2;   // the two warps of a quarter interleave the 32-column chunks
      const int r_in = q * 32 + lane;
      const long long row = (long long)rb * AP_BM + r_in;
      const bool rok = row < sp.rows;
      const float rsum = rs_buf[r_in] + rs_buf[AP_BM + r_in];
      mbar_wait(smem_u32(accfull), aphase);
      aphase ^= 1;
      tc_fence_after();
      const int nch = (sp.dv16 + 31) / 32;
      const bool partial = sp.S > 1;
      float inv = 1.f;
      if (sp.mode == 1) inv = 1.f / (rsum == 0.f ? 1e-30f : rsum);  // marg == 0 -> 1e-30 (eval/match.py:203-204)
      const long long rows_pad = (long long)sp.nblocks * AP_BM;
      for (int ci = grp; ci < nch; ci += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ci * 32, v);
        tmem_ld_wait();
        const int c0 = ci * 32;
        if (partial) {
          float* dst = sp.part + ((long long)split * rows_pad + row) * sp.ldp + c0;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                              __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        } else if (rok) {
          float* dst = sp.Z + row * sp.ldz + c0;
          const float* wsrc = sp.mode == 2 ? sp.W + row * sp.ldw + c0 : nullptr;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (c0 + e < sp.dv) {
              const float acc = __uint_as_float(v[e]);
              dst[e] = sp.mode == 2 ? sp.scale * (rsum * wsrc[e] - acc) : acc * inv;
            }
          }
        }
      }
      if (grp == 0) {
        if (partial)
          sp.part_rs[(long long)split * rows_pad + row] = rsum;
        else if (rok && sp.rowsum_out)
          sp.rowsum_out[row] = rsum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(accempty));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// fold the K-split partials in fixed order and apply the epilogue
__global__ void __launch_bounds__(256) apply_tc_finish_kernel(const ApplySub sp) {
  const long long rows_pad = (long long)sp.nblocks * AP_BM;
  const long long total = (long long)sp.rows * sp.dv;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long row = idx / sp.dv;
    const int c = (int)(idx - row * sp.dv);
    float acc = 0.f, rs = 0.f;
    for (int s = 0; s < sp.S; ++s) {
      acc += sp.part[((long long)s * rows_pad + row) * sp.ldp + c];
      rs += sp.part_rs[(long long)s * rows_pad + row];
    }
    float out;
    if (sp.mode == 2)
      out = sp.scale * (rs * sp.W[row * sp.ldw + c] - acc);
    else if (sp.mode == 1)
      out = acc / (rs == 0.f ? 1e-30f : rs);
    else
      out = acc;
    sp.Z[row * sp.ldz + c] = out;
    if (c == 0 && sp.rowsum_out) sp.rowsum_out[row] = rs;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct ApplyPlanCfg {
  int dv16, TR, nhalves, ldp, nkb, nblocks, S, kb_per_split;
  size_t vt_bytes, part_bytes, rs_bytes, total;
};

static int apply_pick_split(int nblocks, int nkb, int nsm) {
  // cost model in units of one K step: an item costs its K steps plus ~48 steps of fixed work (pipeline fill,
  // epilogue, partial write-out); the launch takes ceil(items / SMs) rounds
  int best = 1;
  double best_t = 1e300;
  for (int S = 1; S <= 16; S *= 2) {
    const int kbps = (nkb + S - 1) / S;
    if (S > 1 && kbps < 16) break;
    const long long items = (long long)nblocks * S;
    const double rounds = (double)((items + nsm - 1) / nsm);
    const double t = rounds * (kbps + 48.0) + (S > 1 ? 24.0 : 0.0);
    if (t < best_t * 0.97) {  // prefer fewer splits unless clearly better
      best_t = t;
      best = S;
    }
  }
  return best;
}

static ApplyPlanCfg apply_cfg(int rows, int kdim, int dv) {
  ApplyPlanCfg c;
  c.dv16 = (dv + 15) / 16 * 16;
  c.nhalves = c.dv16 > AP_NH ? 2 : 1;
  c.TR = c.nhalves == 2 ? AP_NH : c.dv16;
  c.ldp = (c.dv16 + 31) / 32 * 32;
  c.nkb = (kdim + AP_KC - 1) / AP_KC;
  c.nblocks = (rows + AP_BM - 1) / AP_BM;
  c.S = apply_pick_split(c.nblocks, c.nkb, sm_count());
  c.kb_per_split = (c.nkb + c.S - 1) / c.S;
  c.S = (c.nkb + c.kb_per_split - 1) / c.kb_per_split;  // no empty splits
  c.vt_bytes = ((size_t)c.nkb * AP_PARTS * c.nhalves * c.TR * AP_KC * 2 + 1023) / 1024 * 1024;
  const size_t rows_pad = (size_t)c.nblocks * AP_BM;
  c.part_bytes = c.S > 1 ? ((size_t)c.S * rows_pad * c.ldp * 4 + 1023) / 1024 * 1024 : 0;
  c.rs_bytes = c.S > 1 ? ((size_t)c.S * rows_pad * 4 + 1023) / 1024 * 1024 : 0;
  c.total = c.vt_bytes + c.part_bytes + c.rs_bytes;
  return c;
}

static size_t apply_smem(const ApplyPlanCfg& a, const ApplyPlanCfg* b) {
  size_t vstage = (size_t)AP_PARTS * a.nhalves * a.TR * AP_KC * 2;
  if (b) {
    const size_t v1 = (size_t)AP_PARTS * b->nhalves * b->TR * AP_KC * 2;
    vstage = v1 > vstage ? v1 : vstage;
  }
  return (size_t)AP_STAGES * AP_PARTS * AP_A_TILE + AP_STAGES * vstage + 2 * 2 * AP_BM * 4 + (3 * AP_STAGES + 2) * 8 + 16;
}

struct ApplyWeights {  // W_ij = P_ij (w0 + w1 C_ij + wrow_i + wcol_j); wrow over the rows of C, wcol over its columns
  int on;
  float w0, w1;
  const float* wrow;
  const float* wcol;
};

static int apply_fill_sub(ApplySub& s, const ApplyPlanCfg& c, const float* C, int ldc, int n, int m, const float* f,
                          const float* g, int transpose, int dv, int mode, const float* W, int ldw, float scale,
                          float* Z, int ldz, float* rowsum_out, uint8_t* ws, const ApplyWeights* wt = nullptr) {
  s.weighted = (wt && wt->on) ? 1 : 0;
  s.w0 = s.weighted ? wt->w0 : 1.f;
  s.w1 = s.weighted ? wt->w1 : 0.f;
  s.wrow = s.weighted ? (transpose ? wt->wcol : wt->wrow) : nullptr;
  s.wk = s.weighted ? (transpose ? wt->wrow : wt->wcol) : nullptr;
  s.C = C;
  s.ldc = ldc;
  s.frow = transpose ? g : f;
  s.fk = transpose ? f : g;
  s.Vt = ws;
  s.W = W;
  s.ldw = ldw;
  s.Z = Z;
  s.ldz = ldz;
  s.rowsum_out = rowsum_out;
  s.part = c.S > 1 ? reinterpret_cast<float*>(ws + c.vt_bytes) : nullptr;
  s.part_rs = c.S > 1 ? reinterpret_cast<float*>(ws + c.vt_bytes + c.part_bytes) : nullptr;
  s.rows = transpose ? m : n;
  s.kdim = transpose ? n : m;
  s.transposed = transpose;
  s.nkb = c.nkb;
  s.dv = dv;
  s.dv16 = c.dv16;
  s.TR = c.TR;
  s.nhalves = c.nhalves;
  s.ldp = c.ldp;
  s.S = c.S;
  s.kb_per_split = c.kb_per_split;
  s.nblocks = c.nblocks;
  s.mode = mode;
  s.scale = scale;
  // 16-byte loads along the rows of C (both forms); the forward form also reads its K-side potential as float4
  s.vec = ((ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
           (transpose || (reinterpret_cast<uintptr_t>(s.fk) & 15) == 0))
              ? 1
              : 0;
  return 0;
}

static int apply_launch(ApplyArgs& a, const ApplyPlanCfg& c0, const ApplyPlanCfg* c1, cudaStream_t st) {
  const size_t smem = apply_smem(c0, c1);
  B200OT_CUDA_OK(cudaFuncSetAttribute(apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int total = a.items[0] + (a.nsub > 1 ? a.items[1] : 0);
  int grid = sm_count();
  if (grid > total) grid = total;
  apply_tc_kernel<<<grid, AP_THREADS, smem, st>>>(a);
  B200OT_LAUNCH_OK();
  for (int i = 0; i < a.nsub; ++i) {
    if (a.sub[i].S > 1) {
      const long long tot = (long long)a.sub[i].rows * a.sub[i].dv;
      long long blocks = (tot + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      apply_tc_finish_kernel<<<(int)blocks, 256, 0, st>>>(a.sub[i]);
      B200OT_LAUNCH_OK();
    }
  }
  return 0;
}

static int apply_split_v(const float* V, int ldv, int rows_in, int dv, const ApplyPlanCfg& c, uint8_t* out,
                         cudaStream_t st) {
  const long long units = (long long)c.nkb * 4 * c.nhalves * c.TR;
  long long blocks = (units + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  apply_split_v_kernel<<<(int)blocks, 256, 0, st>>>(V, ldv, rows_in, dv, c.TR, c.nhalves, c.nkb, out);
  B200OT_LAUNCH_OK();
  return 0;
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

size_t b200ot_apply_plan_tc_workspace_bytes(int n, int m, int dv, int transpose) {
  if (n <= 0 || m <= 0 || dv <= 0 || dv > AP_MAX_DV) return 0;
  return apply_cfg(transpose ? m : n, transpose ? n : m, dv).total;
}

int b200ot_apply_plan_tc(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                         const float* V, int ldv, int dv, int transpose, int normalise, float* Z, int ldz,
                         float* rowsum_out, void* ws, size_t ws_bytes, void* stream) {
  if (!C || !f || !g || !V || !Z || !ws || n <= 0 || m <= 0 || dv <= 0 || ldc < m || ldv < dv || ldz < dv ||
      !(eps > 0.f))
    return B200OT_E_INVALID;
  if (dv > AP_MAX_DV) return B200OT_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const ApplyPlanCfg c = apply_cfg(transpose ? m : n, transpose ? n : m, dv);
  if (ws_bytes < c.total) return B200OT_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = static_cast<uint8_t*>(ws);
  int rc = apply_split_v(V, ldv, transpose ? n : m, dv, c, w, st);
  if (rc) return rc;
  ApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.nsub = 1;
  a.k = kLog2e / eps;
  apply_fill_sub(a.sub[0], c, C, ldc, n, m, f, g, transpose ? 1 : 0, dv, normalise ? 1 : 0, nullptr, 0, 1.f, Z, ldz,
                 rowsum_out, w);
  a.items[0] = c.nblocks * c.S;
  return apply_launch(a, c, nullptr, st);
}

int b200ot_apply_plan_tc_weighted(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                                  float w0, float w1, const float* wrow, const float* wcol, const float* V, int ldv,
                                  int dv, int transpose, float* Z, int ldz, float* rowsum_out, void* ws,
                                  size_t ws_bytes, void* stream) {
  if (!C || !f || !g || !V || !Z || !ws || n <= 0 || m <= 0 || dv <= 0 || ldc < m || ldv < dv || ldz < dv ||
      !(eps > 0.f))
    return B200OT_E_INVALID;
  if (dv > AP_MAX_DV) return B200OT_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const ApplyPlanCfg c = apply_cfg(transpose ? m : n, transpose ? n : m, dv);
  if (ws_bytes < c.total) return B200OT_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w = static_cast<uint8_t*>(ws);
  int rc = apply_split_v(V, ldv, transpose ? n : m, dv, c, w, st);
  if (rc) return rc;
  ApplyWeights wt = {1, w0, w1, wrow, wcol};
  ApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.nsub = 1;
  a.k = kLog2e / eps;
  apply_fill_sub(a.sub[0], c, C, ldc, n, m, f, g, transpose ? 1 : 0, dv, 0, nullptr, 0, 1.f, Z, ldz, rowsum_out, w, &wt);
  a.items[0] = c.nblocks * c.S;
  return apply_launch(a, c, nullptr, st);
}

int b200ot_envelope_bwd_weighted(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                                 float w0, float w1, const float* wrow, const float* wcol, const float* X, int ldx,
                                 const float* Y, int ldy, int d, float scale, float* dX, int lddx, float* dY, int lddy,
                                 void* ws, size_t ws_bytes, void* stream) {
  if (!C || !f || !g || !X || !Y || !dX || !dY || !ws || n <= 0 || m <= 0 || d <= 0 || ldc < m || ldx < d ||
      ldy < d || lddx < d || lddy < d || !(eps > 0.f))
    return B200OT_E_INVALID;
  if (d > AP_MAX_DV) return B200OT_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const ApplyPlanCfg c0 = apply_cfg(n, m, d);
  const ApplyPlanCfg c1 = apply_cfg(m, n, d);
  if (ws_bytes < c0.total + c1.total) return B200OT_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w0p = static_cast<uint8_t*>(ws);
  uint8_t* w1p = w0p + c0.total;
  int rc = apply_split_v(Y, ldy, m, d, c0, w0p, st);
  if (rc) return rc;
  rc = apply_split_v(X, ldx, n, d, c1, w1p, st);
  if (rc) return rc;
  ApplyWeights wt = {1, w0, w1, wrow, wcol};
  ApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.nsub = 2;
  a.k = kLog2e / eps;
  apply_fill_sub(a.sub[0], c0, C, ldc, n, m, f, g, 0, d, 2, X, ldx, scale, dX, lddx, nullptr, w0p, &wt);
  apply_fill_sub(a.sub[1], c1, C, ldc, n, m, f, g, 1, d, 2, Y, ldy, scale, dY, lddy, nullptr, w1p, &wt);
  a.items[0] = c0.nblocks * c0.S;
  a.items[1] = c1.nblocks * c1.S;
  return apply_launch(a, c0, &c1, st);
}

size_t b200ot_envelope_bwd_workspace_bytes(int n, int m, int d) {
  if (n <= 0 || m <= 0 || d <= 0 || d > AP_MAX_DV) return 0;
  return apply_cfg(n, m, d).total + apply_cfg(m, n, d).total;
}

int b200ot_envelope_bwd(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                        const float* X, int ldx, const float* Y, int ldy, int d, float scale, float* dX, int lddx,
                        float* dY, int lddy, float* rowsum_out, float* colsum_out, void* ws, size_t ws_bytes,
                        void* stream) {
  if (!C || !f || !g || !X || !Y || !dX || !dY || !ws || n <= 0 || m <= 0 || d <= 0 || ldc < m || ldx < d ||
      ldy < d || lddx < d || lddy < d || !(eps > 0.f))
    return B200OT_E_INVALID;
  if (d > AP_MAX_DV) return B200OT_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const ApplyPlanCfg c0 = apply_cfg(n, m, d);  // dX: rows i, contraction over j with Y
  const ApplyPlanCfg c1 = apply_cfg(m, n, d);  // dY: rows j, contraction over i with X
  if (ws_bytes < c0.total + c1.total) return B200OT_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* w0 = static_cast<uint8_t*>(ws);
  uint8_t* w1 = w0 + c0.total;
  int rc = apply_split_v(Y, ldy, m, d, c0, w0, st);
  if (rc) return rc;
  rc = apply_split_v(X, ldx, n, d, c1, w1, st);
  if (rc) return rc;
  ApplyArgs a;
  memset(&a, 0, sizeof(a));
  a.nsub = 2;
  a.k = kLog2e / eps;
  apply_fill_sub(a.sub[0], c0, C, ldc, n, m, f, g, 0, d, 2, X, ldx, scale, dX, lddx, rowsum_out, w0);
  apply_fill_sub(a.sub[1], c1, C, ldc, n, m, f, g, 1, d, 2, Y, ldy, scale, dY, lddy, colsum_out, w1);
  a.items[0] = c0.nblocks * c0.S;
  a.items[1] = c1.nblocks * c1.S;
  return apply_launch(a, c0, &c1, st);
}

}  // extern "C"
