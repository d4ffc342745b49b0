// Cost construction on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
//   C_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j      or      1 - cos(x_i, y_j)
//
// is the one dense contraction of the OT path (ot.dist at MRI_PET_OT_nojax.py:70-71; the T = I
// case of the feature cost at :121-136).  A plan entry is exp((f+g-C)/eps), so an absolute
// error d in C is a relative error d/eps in the plan: bf16 or tf32 inputs (d ~ 1e-3) are
// unusable.  The contraction is therefore run as a 3-term split product on bf16 tensor cores,
//     x = x1 + x2 + x3 (bf16 parts),
//     terms = 3:  x.y ~= x1.y2 + x2.y1 + x1.y1                          (error ~2^-17 per product)
//     terms = 6:  ... + x1.y3 + x3.y1 + x2.y2                           (error ~2^-24: fp32 grade)
// which is an ordinary bf16 GEMM with K' = terms * K over concatenated parts.
// fp16 mode (B200OT_TERMS_F16_3 / _4): every row is scaled by a power of two so that its largest entry lies in
// [2^9, 2^10) and split into two fp16 parts x = x1 + x2 (11 + 11 significand bits, |x2| <= 2^-11 |x|max; the scaling
// keeps x2 out of the fp16 subnormals), 3 products x1.y2 + x2.y1 + x1.y1 (+ x2.y2): representation error
// <= 2^-23 of the row maximum per entry, i.e. fp32-grade relative to |x||y|, at HALF the tensor work of the 6-term
// bf16 split.  The accumulator is rescaled by 2^-(s_i + t_j) in the epilogue.
// A pre-pass splits the fp32 embeddings and writes the parts *pre-tiled*: every
// (ROWS x 64) operand tile is one contiguous block already in the canonical no-swizzle K-major
// UMMA shared-memory layout (8 x 16-byte core matrices), so the GEMM feeds shared memory with
// plain 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) and needs no tensor map.
//
// GEMM kernel: persistent, one CTA per SM, warp-specialised:
//   warp 0  TMA producer     (one lane): ring of 4 stages x (A 128x64 + B 256x64) bf16
//   warp 1  MMA issuer       (one lane): tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16,
//                                        fp32 accumulators in TMEM, two accumulator stages
//   warps 2-5 epilogue       tcgen05.ld -> |x|^2 + |y|^2 - 2 acc -> smem transpose -> coalesced stores
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200ot {

constexpr int TC_BM = 128;   // tile rows (UMMA M)
constexpr int TC_BN = 256;   // tile cols (UMMA N)
constexpr int TC_BK = 64;    // K elements per stage (bf16: 128 B per row)
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_STRIDE = 36;  // floats per staged row of a 32 x 32 chunk: 16-byte accesses stay conflict-free
constexpr int TC_EPI_FLOATS = 32 * TC_EPI_STRIDE + 2 * 256;  // per epilogue warp: chunk staging + column terms of a tile
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KiB
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 2;  // 32 KiB
constexpr int TC_GROUP_N = 16;  // n-blocks per raster group (keeps a 12 MiB slab of B' hot in L2)

// ---- pre-pass: fp32 rows -> two bf16 parts in tiled UMMA layout + row norms ---------------------
// One warp per (padded) row.  x = p1 + p2 + p3 (bf16 parts, 8 + 8 + 8 significand bits).
// Tile (rb, part, kblk) of `tile_rows` x 64 bf16 starts at
// ((rb * nparts + part) * kblocks + kblk) * tile_rows * 128 bytes; inside it the core matrix
// (row group rg, k chunk kc) sits at (kc * tile_rows/8 + rg) * 128 and row r%8 at r%8 * 16.
__global__ void __launch_bounds__(256) split_tiles_kernel(const float* __restrict__ X, long long ldx, int rows,
                                                          int d, int tile_rows, int kblocks, int rows_pad,
                                                          int cosine, int nparts, uint8_t* __restrict__ out,
                                                          float* __restrict__ norms, float prescale, int f16) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows_pad) return;
  const bool live = row < rows;
  const float* x = X + (long long)row * ldx;
  float ss = 0.f, mx = 0.f;
  if (live)
    for (int k = lane; k < d; k += 32) {
      ss = fmaf(x[k], x[k], ss);
      mx = fmaxf(mx, fabsf(x[k]));
    }
  ss = warp_sum(ss);
  float scale = cosine ? (ss > 0.f ? rsqrtf(ss) : 0.f) : prescale;  // prescale: a power of two (exact)
  if (f16) {
    // fp16 mode: norms holds {norm, 2^-s} per row; s puts the row maximum into [2^9, 2^10)
    mx = warp_max(mx) * fabsf(scale);
    const int sexp = (mx > 0.f && mx < INFINITY) ? 9 - ilogbf(mx) : 0;
    if (live && lane == 0) {
      norms[2 * row] = cosine ? 1.f : ss;
      norms[2 * row + 1] = ldexpf(1.f, -sexp);
    }
    scale *= ldexpf(1.f, sexp);
  } else if (live && lane == 0) {
    norms[row] = cosine ? 1.f : ss;
  }
  const int rb = row / tile_rows, rin = row % tile_rows;
  const size_t tile_bytes = (size_t)tile_rows * 128;
  const int nchunks = kblocks * 8;
  for (int c = lane; c < nchunks; c += 32) {
    const int kblk = c >> 3, kc = c & 7;
    __nv_bfloat16 p1[8], p2[8], p3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = c * 8 + e;
      const float v = (live && k < d) ? x[k] * scale : 0.f;
      if (f16) {  // same 16-bit containers, fp16 bit patterns
        const __half h = __float2half_rn(v);
        const __half h2 = __float2half_rn(v - __half2float(h));  // the residual is exact in fp32
        p1[e] = __ushort_as_bfloat16(__half_as_ushort(h));
        p2[e] = __ushort_as_bfloat16(__half_as_ushort(h2));
        p3[e] = __ushort_as_bfloat16((unsigned short)0);
        continue;
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(h);  // exact in fp32
      const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
      p1[e] = h;
      p2[e] = h2;
      p3[e] = __float2bfloat16_rn(r1 - __bfloat162float(h2));
    }
    const size_t inner = ((size_t)kc * (tile_rows / 8) + (rin >> 3)) * 128 + (size_t)(rin & 7) * 16;
    const size_t tb = (size_t)rb * nparts * kblocks + kblk;
    *reinterpret_cast<uint4*>(out + tb * tile_bytes + inner) = *reinterpret_cast<const uint4*>(p1);
    if (nparts > 1)
      *reinterpret_cast<uint4*>(out + (tb + kblocks) * tile_bytes + inner) = *reinterpret_cast<const uint4*>(p2);
    if (nparts > 2)
      *reinterpret_cast<uint4*>(out + (tb + 2 * (size_t)kblocks) * tile_bytes + inner) =
          *reinterpret_cast<const uint4*>(p3);
  }
}

// ---- the GEMM -----------------------------------------------------------------------------------
struct CostTcArgs {
  const uint8_t* A;   // tiled parts of X  (tile_rows = TC_BM)
  const uint8_t* B;   // tiled parts of Y  (tile_rows = TC_BN)
  const float* xn;    // |x_i|^2 (or 1)
  const float* yn;
  float* C;
  long long ldc;
  int n, m, kblocks, nseg, nparts, cosine;
  int tiles_m, tiles_n;
  int f16;  // fp16 parts: xn / yn hold {norm, 2^-s} per row, the accumulator is rescaled in the epilogue
  int vec_ok;  // C is 16-byte aligned with ldc % 4 == 0: rows are stored as float4
};

// segment -> (part of X, part of Y): smallest cross terms first, the x1.y1 term last
__device__ __forceinline__ void tc_segment(int nseg, int seg, int& pa, int& pb) {
  pa = 0;
  pb = 0;
  if (nseg == 3) {  // x1.y2, x2.y1, x1.y1
    pa = seg == 1 ? 1 : 0;
    pb = seg == 0 ? 1 : 0;
  } else if (nseg == 4) {  // x2.y2, x1.y2, x2.y1, x1.y1
    pa = (seg == 0 || seg == 2) ? 1 : 0;
    pb = (seg == 0 || seg == 1) ? 1 : 0;
  } else if (nseg == 6) {  // x1.y3, x3.y1, x2.y2, x1.y2, x2.y1, x1.y1
    pa = seg == 1 ? 2 : ((seg == 2 || seg == 4) ? 1 : 0);
    pb = seg == 0 ? 2 : ((seg == 2 || seg == 3) ? 1 : 0);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) cost_tc_kernel(const CostTcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint8_t* sA = smem;                                    // TC_STAGES x 16 KiB
  uint8_t* sB = smem + TC_STAGES * TC_A_BYTES;           // TC_STAGES x 32 KiB
  float* sEpi = reinterpret_cast<float*>(smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));  // per-warp epilogue scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + 4 * TC_EPI_FLOATS);
  uint64_t* full = bars;                 // [TC_STAGES]
  uint64_t* empty = bars + TC_STAGES;    // [TC_STAGES]
  uint64_t* tfull = bars + 2 * TC_STAGES;      // [2]
  uint64_t* tempty = bars + 2 * TC_STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total_tiles = (long long)p.tiles_m * p.tiles_n;
  const int ksteps = p.kblocks * p.nseg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(tfull + a), 1);
      mbar_init(smem_u32(tempty + a), 4);
    }
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

  auto tile_coords = [&](long long t, int& mb, int& nb) {
    const long long per_group = (long long)p.tiles_m * TC_GROUP_N;
    const int g = (int)(t / per_group);
    const long long idx = t - (long long)g * per_group;
    int gw = p.tiles_n - g * TC_GROUP_N;
    gw = gw > TC_GROUP_N ? TC_GROUP_N : gw;
    mb = (int)(idx / gw);
    nb = g * TC_GROUP_N + (int)(idx - (long long)mb * gw);
  };
  // the last raster group may be narrower than TC_GROUP_N; per_group above assumes full width, so
  // tiles are enumerated group by group with the true width:
  auto tile_coords_exact = [&](long long t, int& mb, int& nb) {
    const long long full_groups = p.tiles_n / TC_GROUP_N;
    const long long per_group = (long long)p.tiles_m * TC_GROUP_N;
    if (t < full_groups * per_group) {
      tile_coords(t, mb, nb);
    } else {
      const long long idx = t - full_groups * per_group;
      const int gw = p.tiles_n - (int)full_groups * TC_GROUP_N;
      mb = (int)(idx / gw);
      nb = (int)full_groups * TC_GROUP_N + (int)(idx - (long long)mb * gw);
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t a_tile = TC_A_BYTES, b_tile = TC_B_BYTES;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int mb, nb;
        tile_coords_exact(t, mb, nb);
        for (int ks = 0; ks < ksteps; ++ks) {
          const int seg = ks / p.kblocks, kblk = ks - seg * p.kblocks;
          // segment -> (part of X, part of Y): smallest cross terms first, the x1.y1 term last
          int pa, pb;
          tc_segment(p.nseg, seg, pa, pb);
          mbar_wait(smem_u32(empty + stage), phase ^ 1);
          const uint32_t bar = smem_u32(full + stage);
          mbar_arrive_expect_tx(bar, TC_A_BYTES + TC_B_BYTES);
          const uint8_t* srcA = p.A + ((size_t)(mb * p.nparts + pa) * p.kblocks + kblk) * a_tile;
          const uint8_t* srcB = p.B + ((size_t)(nb * p.nparts + pb) * p.kblocks + kblk) * b_tile;
          bulk_g2s(smem_u32(sA + (size_t)stage * TC_A_BYTES), srcA, TC_A_BYTES, bar);
          bulk_g2s(smem_u32(sB + (size_t)stage * TC_B_BYTES), srcB, TC_B_BYTES, bar);
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
      // (fp16 parts: A / B format fields 0 instead of 1)
      const uint32_t idesc = (1u << 4) | (p.f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(TC_BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(smem_u32(tempty + as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * TC_BN;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(smem_u32(full + stage), phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)stage * TC_A_BYTES);
          const uint32_t b0 = smem_u32(sB + (size_t)stage * TC_B_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // one MMA consumes two 8-element K chunks: chunk stride = tile_rows * 16 bytes
            const uint64_t ad = make_smem_desc(a0 + (uint32_t)k * 2u * (TC_BM * 16u), TC_BM * 16u, 128u);
            const uint64_t bd = make_smem_desc(b0 + (uint32_t)k * 2u * (TC_BN * 16u), TC_BN * 16u, 128u);
            tc_mma_bf16(tmem_d, ad, bd, idesc, (ks > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(empty + stage));  // frees the smem stage once these MMAs have read it
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(smem_u32(tfull + as));  // accumulator complete
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    float* sm = sEpi + (warp - 2) * TC_EPI_FLOATS;  // [32][36] chunk staging
    float* sy = sm + 32 * TC_EPI_STRIDE;            // [2][256] column terms of the tile
    int as = 0;
    uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int mb, nb;
      tile_coords_exact(t, mb, nb);
      // tcgen05.ld 32x32b hands lane l the 32 consecutive columns of row (quarter * 32 + l).  Every lane finishes
      // its own row (row terms are its own registers; the 256 column terms of the tile are staged in shared memory
      // while the tile's MMAs are still running and read back as broadcast float4) and writes the 32 results into a
      // [32][36] staging block; the warp then reads the block back with 8 lanes per row and stores 4 full 128-byte
      // lines per instruction.  The TMEM load of the next chunk is in flight while a chunk is finished.
      // History (profiles/r02_cost_tc_f16_16384_*.txt): (1) transposing scalar by scalar with two shuffles, 64-bit
      // address arithmetic and a branch per row and chunk cost ~30 instructions per row: 20 us per tile, tensor pipe
      // capped at 37 % with three K segments; (2) storing 16 bytes per lane straight from the TMEM registers needs
      // no staging but every store instruction touches 32 half-filled sectors: 14 us per tile, 59 %.
      const int row = mb * TC_BM + q * 32 + lane;
      const bool rok = row < p.n;
      float xn_l = 0.f, xs_l = 1.f;
      if (rok) {
        xn_l = p.f16 ? p.xn[2 * row] : p.xn[row];
        if (p.f16) xs_l = p.xn[2 * row + 1];
      }
      __syncwarp();  // the previous tile's readers are done with the scratch
#pragma unroll
      for (int c = 0; c < TC_BN / 32; ++c) {
        const int col = nb * TC_BN + c * 32 + lane;
        const bool cok = col < p.m;
        sy[c * 32 + lane] = (cok && !p.cosine) ? (p.f16 ? p.yn[2 * col] : p.yn[col]) : 0.f;
        sy[TC_BN + c * 32 + lane] = (cok && p.f16) ? p.yn[2 * col + 1] : 1.f;
      }
      __syncwarp();
      const float mul_l = p.cosine ? -xs_l : -2.f * xs_l;  // powers of two times -1 / -2: exact
      const float base_l = p.cosine ? 1.f : xn_l;
      // read-back role of this lane: row (lane / 8) + 4 i of the chunk, columns (lane % 8) * 4 .. + 3
      const int rb_row = lane >> 3, rb_col = (lane & 7) * 4;
      const int row_q0 = mb * TC_BM + q * 32;
      const bool tile_fast = p.vec_ok && row_q0 + 32 <= p.n;  // full rows, aligned: the coalesced path
      mbar_wait(smem_u32(tfull + as), aphase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * TC_BN;
      auto finish = [&](int c, const uint32_t (&v)[32]) {
        const int col0 = nb * TC_BN + c * 32;
        const float* syc = sy + c * 32;
        float o[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 y4 = *reinterpret_cast<const float4*>(syc + 4 * j4);
          float4 m4 = make_float4(mul_l, mul_l, mul_l, mul_l);
          if (p.f16) {
            const float4 s4 = *reinterpret_cast<const float4*>(syc + TC_BN + 4 * j4);
            m4 = make_float4(mul_l * s4.x, mul_l * s4.y, mul_l * s4.z, mul_l * s4.w);
          }
          o[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), m4.x, base_l + y4.x);
          o[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), m4.y, base_l + y4.y);
          o[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), m4.z, base_l + y4.z);
          o[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), m4.w, base_l + y4.w);
        }
        if (tile_fast && col0 + 32 <= p.m) {
          __syncwarp();  // the previous chunk has been read back
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(sm + lane * TC_EPI_STRIDE + 4 * j4) =
                make_float4(o[4 * j4 + 0], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
          __syncwarp();
          float* dst = p.C + (long long)(row_q0 + rb_row) * p.ldc + col0 + rb_col;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w4 = *reinterpret_cast<const float4*>(sm + (rb_row + 4 * i) * TC_EPI_STRIDE + rb_col);
            *reinterpret_cast<float4*>(dst) = w4;
            dst += 4 * p.ldc;
          }
        } else if (rok) {
          float* crow = p.C + (long long)row * p.ldc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.m) crow[col0 + j] = o[j];
        }
      };
      uint32_t va[32], vb[32];
      tmem_ld32(tacc, va);
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; c += 2) {
        tmem_ld_wait();
        tmem_ld32(tacc + (uint32_t)(c + 1) * 32, vb);
        finish(c, va);
        tmem_ld_wait();
        if (c + 2 < TC_BN / 32) tmem_ld32(tacc + (uint32_t)(c + 2) * 32, va);
        finish(c + 1, vb);
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(smem_u32(tempty + as));
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

constexpr size_t TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 4 * TC_EPI_FLOATS * 4 + (2 * TC_STAGES + 4) * 8 + 16;

struct CostWs {
  size_t a_off, b_off, norm_off, total;
  int kblocks, n_pad, m_pad;
};
static CostWs cost_ws(int n, int m, int d) {
  CostWs w;
  w.kblocks = (d + TC_BK - 1) / TC_BK;
  w.n_pad = (n + TC_BM - 1) / TC_BM * TC_BM;
  w.m_pad = (m + TC_BN - 1) / TC_BN * TC_BN;
  const size_t a_bytes = (size_t)w.n_pad * w.kblocks * TC_BK * 2 * 3;  // up to three bf16 parts
  const size_t b_bytes = (size_t)w.m_pad * w.kblocks * TC_BK * 2 * 3;
  w.a_off = 0;
  w.b_off = (a_bytes + 1023) / 1024 * 1024;
  w.norm_off = w.b_off + (b_bytes + 1023) / 1024 * 1024;
  w.total = w.norm_off + ((size_t)(n + m) * 8 + 1023) / 1024 * 1024;  // {norm, 2^-s} per row in fp16 mode
  return w;
}

// out (cols x rows) = X^T for a rows x cols row-major matrix (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) fot_prep_kernel(const float* __restrict__ X, long long ldx, int rows, int cols,
                                                       float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int row = r0 + r, col = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (row < rows && col < cols) ? X[(long long)row * ldx + col] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += 8) {
    const int col = c0 + c, row = r0 + threadIdx.x;
    if (col < cols && row < rows) out[(long long)col * rows + row] = tile[threadIdx.x][c];
  }
}

// out[k] = sum_i XT(k, i)^2 w_i for the transposed matrix XT (feat x samples, contiguous): one warp per feature
__global__ void __launch_bounds__(256) fot_wnorm_kernel(const float* __restrict__ XT, int samples, int feat,
                                                        const float* __restrict__ w, float* __restrict__ out) {
  const int k = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= feat) return;
  float acc = 0.f;
  for (int i = lane; i < samples; i += 32) {
    const float v = XT[(long long)k * samples + i];
    acc = fmaf(v * v, w[i], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[k] = acc;
}

// `terms` of the C ABI -> K segments, parts per operand, fp16 parts
struct TermCfg {
  int nseg, nparts, f16;
};
static bool decode_terms(int terms, TermCfg* c) {
  switch (terms) {
    case 1: *c = {1, 1, 0}; return true;
    case 3: *c = {3, 2, 0}; return true;
    case 6: *c = {6, 3, 0}; return true;
    case B200OT_TERMS_F16_3: *c = {3, 2, 1}; return true;
    case B200OT_TERMS_F16_4: *c = {4, 2, 1}; return true;
    default: return false;
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

size_t b200ot_cost_workspace_bytes(int n, int m, int d) {
  if (n <= 0 || m <= 0 || d <= 0) return 0;
  return cost_ws(n, m, d).total;
}

int b200ot_cost(const float* X, int ldx, const float* Y, int ldy, int n, int m, int d, int kind,
                float* C, int ldc, void* ws, size_t ws_bytes, int terms, void* stream) {
  if (!X || !Y || !C || !ws || n <= 0 || m <= 0 || d <= 0 || ldx < d || ldy < d || ldc < m)
    return B200OT_E_INVALID;
  if (kind != B200OT_COST_SQEUCLIDEAN && kind != B200OT_COST_COSINE) return B200OT_E_INVALID;
  TermCfg tc;
  if (!decode_terms(terms, &tc)) return B200OT_E_INVALID;
  const int nparts = tc.nparts;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const CostWs w = cost_ws(n, m, d);
  if (ws_bytes < w.total) return B200OT_E_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(ws);
  uint8_t* Ap = base + w.a_off;
  uint8_t* Bp = base + w.b_off;
  float* xn = reinterpret_cast<float*>(base + w.norm_off);
  float* yn = xn + (tc.f16 ? 2 * (size_t)n : (size_t)n);
  const int cosine = kind == B200OT_COST_COSINE ? 1 : 0;
  split_tiles_kernel<<<(w.n_pad * 32 + 255) / 256, 256, 0, s>>>(X, ldx, n, d, TC_BM, w.kblocks, w.n_pad, cosine, nparts, Ap, xn, 1.f, tc.f16);
  B200OT_LAUNCH_OK();
  split_tiles_kernel<<<(w.m_pad * 32 + 255) / 256, 256, 0, s>>>(Y, ldy, m, d, TC_BN, w.kblocks, w.m_pad, cosine, nparts, Bp, yn, 1.f, tc.f16);
  B200OT_LAUNCH_OK();
  static PerDeviceOnce attr_once;  // function attributes are per device
  if (attr_once.first()) {
    B200OT_CUDA_OK(cudaFuncSetAttribute(cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  }
  CostTcArgs a;
  a.A = Ap;
  a.B = Bp;
  a.xn = xn;
  a.yn = yn;
  a.C = C;
  a.ldc = ldc;
  a.n = n;
  a.m = m;
  a.kblocks = w.kblocks;
  a.nseg = tc.nseg;
  a.nparts = nparts;
  a.cosine = cosine;
  a.f16 = tc.f16;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc & 3) == 0) ? 1 : 0;
  a.tiles_m = w.n_pad / TC_BM;
  a.tiles_n = w.m_pad / TC_BN;
  const long long tiles = (long long)a.tiles_m * a.tiles_n;
  int grid = sm_count();
  if (grid > tiles) grid = (int)tiles;
  cost_tc_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(a);
  B200OT_LAUNCH_OK();
  return 0;
}

// ---- split / gemm halves of b200ot_cost, for callers that keep the bf16 parts resident (online solver) ----
size_t b200ot_cost_parts_bytes(int rows, int d, int side) {
  if (rows <= 0 || d <= 0) return 0;
  const int tile = side ? TC_BN : TC_BM;
  const size_t rows_pad = (size_t)(rows + tile - 1) / tile * tile;
  const size_t kblocks = (size_t)(d + TC_BK - 1) / TC_BK;
  return (rows_pad * kblocks * TC_BK * 2 * 3 + 1023) / 1024 * 1024;
}

int b200ot_cost_split(const float* X, int ldx, int rows, int d, int kind, int terms, int side, void* parts,
                      float* norms, void* stream) {
  if (!X || !parts || !norms || rows <= 0 || d <= 0 || ldx < d) return B200OT_E_INVALID;
  TermCfg tc;
  if (!decode_terms(terms, &tc)) return B200OT_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(parts) & 1023) != 0) return B200OT_E_INVALID;
  const int nparts = tc.nparts;
  const int tile = side ? TC_BN : TC_BM;
  const int rows_pad = (rows + tile - 1) / tile * tile;
  const int kblocks = (d + TC_BK - 1) / TC_BK;
  split_tiles_kernel<<<(rows_pad * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X, ldx, rows, d, tile, kblocks, rows_pad, kind == B200OT_COST_COSINE ? 1 : 0, nparts,
      static_cast<uint8_t*>(parts), norms, 1.f, tc.f16);
  B200OT_LAUNCH_OK();
  return 0;
}

// C (n x m) from resident parts; row_tile0 selects the first 128-row tile of the A parts (n rows from there)
int b200ot_cost_gemm(const void* partsA, const float* normsA, int row_tile0, int n, const void* partsB,
                     const float* normsB, int m, int d, int kind, int terms, float* C, int ldc, void* stream) {
  if (!partsA || !normsA || !partsB || !normsB || !C || n <= 0 || m <= 0 || d <= 0 || ldc < m || row_tile0 < 0)
    return B200OT_E_INVALID;
  TermCfg tc;
  if (!decode_terms(terms, &tc)) return B200OT_E_INVALID;
  static PerDeviceOnce attr_once;  // function attributes are per device
  if (attr_once.first()) {
    B200OT_CUDA_OK(cudaFuncSetAttribute(cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  }
  const int nparts = tc.nparts;
  const int kblocks = (d + TC_BK - 1) / TC_BK;
  CostTcArgs a;
  a.A = static_cast<const uint8_t*>(partsA) + (size_t)row_tile0 * nparts * kblocks * TC_A_BYTES;
  a.B = static_cast<const uint8_t*>(partsB);
  a.xn = normsA + (size_t)row_tile0 * TC_BM * (tc.f16 ? 2 : 1);
  a.yn = normsB;
  a.C = C;
  a.ldc = ldc;
  a.n = n;
  a.m = m;
  a.kblocks = kblocks;
  a.nseg = tc.nseg;
  a.nparts = nparts;
  a.cosine = kind == B200OT_COST_COSINE ? 1 : 0;
  a.f16 = tc.f16;
  a.vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc & 3) == 0) ? 1 : 0;
  a.tiles_m = (n + TC_BM - 1) / TC_BM;
  a.tiles_n = (m + TC_BN - 1) / TC_BN;
  const long long tiles = (long long)a.tiles_m * a.tiles_n;
  int grid = sm_count();
  if (grid > tiles) grid = (int)tiles;
  cost_tc_kernel<<<grid, TC_THREADS, TC_SMEM, static_cast<cudaStream_t>(stream)>>>(a);
  B200OT_LAUNCH_OK();
  return 0;
}

// ---- FOT feature cost as a tcgen05 chain ---------------------------------------------------------
// M = (A.^2)^T w1 (+) (B.^2)^T w2 - 2 A^T Ts B  (MRI_PET_OT_nojax.py:121-136; perturbot/perturbot/match/fot.py:118-128;
// utils.py:125-184) with A n x d, B n2 x d2, Ts n x n2: two contractions on the tensor cores through the same split /
// GEMM pair as b200ot_cost.  G = -2 B^T Ts^T (d2 x n; norms 0), then M = t1 (+) t2 - 2 A^T (-G/2)^T: the factor -1/2
// is folded into the split of G (a power of two, exact).
struct FotWs {
  size_t t1, t2, zeros, scratch, AT, BT, G, pBT, pTs, pAT, pG, total;
};
static FotWs fot_ws(int n, int n2, int d, int d2) {
  FotWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  };
  const size_t rmax = (size_t)((d > d2 ? d : d2) > n ? (d > d2 ? d : d2) : n) + TC_BN;
  w.t1 = take(sizeof(float) * ((size_t)d + TC_BM));
  w.t2 = take(sizeof(float) * ((size_t)d2 + TC_BN));
  w.zeros = take(sizeof(float) * rmax);
  w.scratch = take(sizeof(float) * rmax);
  w.AT = take(sizeof(float) * (size_t)d * n);
  w.BT = take(sizeof(float) * (size_t)d2 * n2);
  w.G = take(sizeof(float) * (size_t)d2 * n);
  w.pBT = take(b200ot_cost_parts_bytes(d2, n2, 0));
  w.pTs = take(b200ot_cost_parts_bytes(n, n2, 1));
  w.pAT = take(b200ot_cost_parts_bytes(d, n, 0));
  w.pG = take(b200ot_cost_parts_bytes(d2, n, 1));
  w.total = off;
  return w;
}

size_t b200ot_fot_cost_tc_workspace_bytes(int n, int n2, int d, int d2) {
  if (n <= 0 || n2 <= 0 || d <= 0 || d2 <= 0) return 0;
  return fot_ws(n, n2, d, d2).total;
}

int b200ot_fot_cost_tc(const float* A, int lda, const float* B, int ldb, const float* Ts, int ldt, const float* w1,
                       const float* w2, int n, int n2, int d, int d2, float* M, int ldm, void* ws, size_t ws_bytes,
                       void* stream) {
  if (!A || !B || !Ts || !w1 || !w2 || !M || !ws || n <= 0 || n2 <= 0 || d <= 0 || d2 <= 0 || lda < d || ldb < d2 ||
      ldt < n2 || ldm < d2)
    return B200OT_E_INVALID;
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return B200OT_E_INVALID;
  const FotWs w = fot_ws(n, n2, d, d2);
  if (ws_bytes < w.total) return B200OT_E_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* t1 = reinterpret_cast<float*>(base + w.t1);
  float* t2 = reinterpret_cast<float*>(base + w.t2);
  float* zeros = reinterpret_cast<float*>(base + w.zeros);
  float* scratch = reinterpret_cast<float*>(base + w.scratch);
  float* AT = reinterpret_cast<float*>(base + w.AT);
  float* BT = reinterpret_cast<float*>(base + w.BT);
  float* G = reinterpret_cast<float*>(base + w.G);
  B200OT_CUDA_OK(cudaMemsetAsync(zeros, 0, w.scratch - w.zeros, s));
  // t1_k = sum_i A(i,k)^2 w1_i, t2_l = sum_j B(j,l)^2 w2_j; transposes (contraction index last) in the same pass
  fot_prep_kernel<<<dim3((d + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, s>>>(A, lda, n, d, AT);
  B200OT_LAUNCH_OK();
  fot_prep_kernel<<<dim3((d2 + 31) / 32, (n2 + 31) / 32), dim3(32, 8), 0, s>>>(B, ldb, n2, d2, BT);
  B200OT_LAUNCH_OK();
  fot_wnorm_kernel<<<(d * 32 + 255) / 256, 256, 0, s>>>(AT, n, d, w1, t1);
  B200OT_LAUNCH_OK();
  fot_wnorm_kernel<<<(d2 * 32 + 255) / 256, 256, 0, s>>>(BT, n2, d2, w2, t2);
  B200OT_LAUNCH_OK();
  const int terms = 6, nparts = 3;
  auto split = [&](const float* X, int ldx, int rows, int K, int side, size_t off, float prescale) {
    const int tile = side ? TC_BN : TC_BM;
    const int rows_pad = (rows + tile - 1) / tile * tile;
    const int kblocks = (K + TC_BK - 1) / TC_BK;
    split_tiles_kernel<<<(rows_pad * 32 + 255) / 256, 256, 0, s>>>(X, ldx, rows, K, tile, kblocks, rows_pad, 0, nparts,
                                                                   reinterpret_cast<uint8_t*>(base + off), scratch,
                                                                   prescale, 0);
  };
  split(BT, n2, d2, n2, 0, w.pBT, 1.f);
  B200OT_LAUNCH_OK();
  split(Ts, ldt, n, n2, 1, w.pTs, 1.f);
  B200OT_LAUNCH_OK();
  int rc = b200ot_cost_gemm(base + w.pBT, zeros, 0, d2, base + w.pTs, zeros, n, n2, B200OT_COST_SQEUCLIDEAN, terms, G,
                            n, stream);
  if (rc) return rc;
  split(AT, n, d, n, 0, w.pAT, 1.f);
  B200OT_LAUNCH_OK();
  split(G, n, d2, n, 1, w.pG, -0.5f);
  B200OT_LAUNCH_OK();
  return b200ot_cost_gemm(base + w.pAT, t1, 0, d, base + w.pG, t2, d2, n, B200OT_COST_SQEUCLIDEAN, terms, M, ldm,
                          stream);
}

}  // extern "C"
