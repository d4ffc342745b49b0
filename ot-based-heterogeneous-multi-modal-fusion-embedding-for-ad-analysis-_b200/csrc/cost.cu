// Cost construction (SIMT fp32 reference-accuracy path) and small dense helpers.
//
//   b200ot_cost_simt : C_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j or 1 - cos(x_i, y_j)
//                      (ot.dist at MRI_PET_OT_nojax.py:70-71; T = I case of :121-136)
//   b200ot_fot_cost  : M = (A.^2)^T w1 (+) (B.^2)^T w2 - 2 A^T Ts B
//                      (MRI_PET_OT_nojax.py:121-136, perturbot/perturbot/match/fot.py:118-128,
//                       perturbot/perturbot/match/utils.py:125-184)
//   b200ot_matrix_max / b200ot_matrix_scale_by_inv : ott Geometry(scale_cost="max_cost")
//                      (perturbot/perturbot/match/fot.py:129-133)
// The tcgen05 version of the big contraction lives in cost_tc.cu.
#include <math.h>

#include "common.cuh"

namespace b200ot {

// ---- generic strided tiled GEMM: D(i,j) = alpha * sum_k A(i,k) B(k,j) + rowv_i + colv_j ----
// A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj].  128x128x16 tiles, 8x8 per thread.
constexpr int GM = 128, GN = 128, GK = 16, GT = 256;

enum GemmEpi { EPI_PLAIN = 0, EPI_SQEUCLID = 1, EPI_COSINE = 2 };

struct GemmArgs {
  const float* A;
  long long sai, sak;
  const float* B;
  long long sbk, sbj;
  float* D;
  long long ldd;
  int M, N, K;
  float alpha;
  const float* rowv;  // EPI_SQEUCLID: |x|^2 ; EPI_COSINE: |x| ; EPI_PLAIN: added if non-null
  const float* colv;
  int epi;
};

__global__ void __launch_bounds__(GT) gemm_strided_kernel(const GemmArgs g) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * GM, j0 = blockIdx.x * GN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 x 8 (strided by 16)
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  const bool a_kfast = (g.sak == 1);
  const bool b_kfast = (g.sbk == 1);
  for (int k0 = 0; k0 < g.K; k0 += GK) {
    // stage A tile (GM x GK) and B tile (GK x GN); the thread->element map follows the unit stride
#pragma unroll
    for (int e = 0; e < (GM * GK) / GT; ++e) {
      const int idx = e * GT + tid;
      int ii, kk;
      if (a_kfast) {
        kk = idx & (GK - 1);
        ii = idx >> 4;
      } else {
        ii = idx & (GM - 1);
        kk = idx >> 7;
      }
      const int gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < g.M && gk < g.K) ? g.A[gi * g.sai + gk * g.sak] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < (GN * GK) / GT; ++e) {
      const int idx = e * GT + tid;
      int jj, kk;
      if (b_kfast) {
        kk = idx & (GK - 1);
        jj = idx >> 4;
      } else {
        jj = idx & (GN - 1);
        kk = idx >> 7;
      }
      const int gj = j0 + jj, gk = k0 + kk;
      Bs[kk][jj] = (gj < g.N && gk < g.K) ? g.B[gk * g.sbk + gj * g.sbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[8], bv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) av[r] = As[kk][ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 8; ++c) bv[c] = Bs[kk][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int gi = i0 + ty + 16 * r;
    if (gi >= g.M) continue;
    const float rv = g.rowv ? g.rowv[gi] : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int gj = j0 + tx + 16 * c;
      if (gj >= g.N) continue;
      const float cv = g.colv ? g.colv[gj] : 0.f;
      float v;
      if (g.epi == EPI_SQEUCLID)
        v = (rv + cv) - 2.f * acc[r][c];
      else if (g.epi == EPI_COSINE)
        v = 1.f - acc[r][c] / fmaxf(rv * cv, 1e-30f);
      else
        v = g.alpha * acc[r][c] + rv + cv;
      g.D[(long long)gi * g.ldd + gj] = v;
    }
  }
}

static int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM);
  gemm_strided_kernel<<<grid, GT, 0, s>>>(g);
  B200OT_LAUNCH_OK();
  return 0;
}

// out[i] = sum_k w_k * X(i,k)^2  (w == nullptr: plain squared norm; sqrt_out: take the root)
__global__ void __launch_bounds__(256) wsqnorm_kernel(const float* __restrict__ X, long long si,
                                                      long long sk, int rows, int K,
                                                      const float* __restrict__ w, float* out,
                                                      int sqrt_out) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = X[warp * si + k * sk];
    s = fmaf(v * v, w ? w[k] : 1.f, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[warp] = sqrt_out ? sqrtf(s) : s;
}

__global__ void __launch_bounds__(256) matrix_max_kernel(const float* __restrict__ C, long long ldc,
                                                         int n, int m, float* out) {
  float mx = -INFINITY;
  for (long long r = blockIdx.x; r < n; r += gridDim.x)
    for (int j = threadIdx.x; j < m; j += 256) mx = fmaxf(mx, C[r * ldc + j]);
  mx = warp_max(mx);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sh[w]);
    // float max via ordered-int atomics (works for mixed signs)
    int* o = reinterpret_cast<int*>(out);
    if (mx >= 0.f)
      atomicMax(o, __float_as_int(mx));
    else
      atomicMin(reinterpret_cast<unsigned int*>(o), __float_as_uint(mx));
  }
}
__global__ void set_float_kernel(float* p, float v) { *p = v; }

__global__ void __launch_bounds__(256) matrix_scale_kernel(float* C, long long ldc, int n, int m,
                                                           const float* denom) {
  const float inv = 1.f / *denom;
  for (long long r = blockIdx.y; r < n; r += gridDim.y)
    for (int j = blockIdx.x * 256 + threadIdx.x; j < m; j += gridDim.x * 256) C[r * ldc + j] *= inv;
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_cost_simt(const float* X, int ldx, const float* Y, int ldy, int n, int m, int d,
                     int kind, float* C, int ldc, float* norms, void* stream) {
  if (!X || !Y || !C || !norms || n <= 0 || m <= 0 || d <= 0 || ldx < d || ldy < d || ldc < m)
    return B200OT_E_INVALID;
  if (kind != B200OT_COST_SQEUCLIDEAN && kind != B200OT_COST_COSINE) return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* xn = norms;
  float* yn = norms + n;
  const int sq = kind == B200OT_COST_COSINE ? 1 : 0;
  wsqnorm_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(X, ldx, 1, n, d, nullptr, xn, sq);
  B200OT_LAUNCH_OK();
  wsqnorm_kernel<<<(m * 32 + 255) / 256, 256, 0, s>>>(Y, ldy, 1, m, d, nullptr, yn, sq);
  B200OT_LAUNCH_OK();
  GemmArgs g;
  g.A = X;
  g.sai = ldx;
  g.sak = 1;
  g.B = Y;
  g.sbk = 1;
  g.sbj = ldy;
  g.D = C;
  g.ldd = ldc;
  g.M = n;
  g.N = m;
  g.K = d;
  g.alpha = 1.f;
  g.rowv = xn;
  g.colv = yn;
  g.epi = kind == B200OT_COST_COSINE ? EPI_COSINE : EPI_SQEUCLID;
  return launch_gemm(g, s);
}

int b200ot_fot_cost(const float* A, int lda, const float* B, int ldb, const float* Ts, int ldt,
                    const float* w1, const float* w2, int n, int n2, int d, int d2, float* M,
                    int ldm, float* tmp, void* stream) {
  if (!A || !B || !Ts || !w1 || !w2 || !M || !tmp || n <= 0 || n2 <= 0 || d <= 0 || d2 <= 0)
    return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tmp layout: [n x d2] Ts.B, then t1 (d), t2 (d2)
  float* TB = tmp;
  float* t1 = tmp + (size_t)n * d2;
  float* t2 = t1 + d;
  // t1_k = sum_i A(i,k)^2 w1_i : "rows" are features k, reduction runs over samples i
  wsqnorm_kernel<<<(d * 32 + 255) / 256, 256, 0, s>>>(A, 1, lda, d, n, w1, t1, 0);
  B200OT_LAUNCH_OK();
  wsqnorm_kernel<<<(d2 * 32 + 255) / 256, 256, 0, s>>>(B, 1, ldb, d2, n2, w2, t2, 0);
  B200OT_LAUNCH_OK();
  GemmArgs g;
  // TB = Ts (n x n2) . B (n2 x d2)
  g.A = Ts;
  g.sai = ldt;
  g.sak = 1;
  g.B = B;
  g.sbk = ldb;
  g.sbj = 1;
  g.D = TB;
  g.ldd = d2;
  g.M = n;
  g.N = d2;
  g.K = n2;
  g.alpha = 1.f;
  g.rowv = nullptr;
  g.colv = nullptr;
  g.epi = EPI_PLAIN;
  int rc = launch_gemm(g, s);
  if (rc) return rc;
  // M = -2 A^T (d x n) . TB (n x d2) + t1 (+) t2
  g.A = A;
  g.sai = 1;
  g.sak = lda;
  g.B = TB;
  g.sbk = d2;
  g.sbj = 1;
  g.D = M;
  g.ldd = ldm;
  g.M = d;
  g.N = d2;
  g.K = n;
  g.alpha = -2.f;
  g.rowv = t1;
  g.colv = t2;
  g.epi = EPI_PLAIN;
  return launch_gemm(g, s);
}

int b200ot_matrix_max(const float* C, int ldc, int n, int m, float* out_max, void* stream) {
  if (!C || !out_max || n <= 0 || m <= 0 || ldc < m) return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  set_float_kernel<<<1, 1, 0, s>>>(out_max, -INFINITY);
  B200OT_LAUNCH_OK();
  const int grid = n < 148 * 8 ? n : 148 * 8;
  matrix_max_kernel<<<grid, 256, 0, s>>>(C, ldc, n, m, out_max);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_matrix_scale_by_inv(float* C, int ldc, int n, int m, const float* denom, void* stream) {
  if (!C || !denom || n <= 0 || m <= 0 || ldc < m) return B200OT_E_INVALID;
  dim3 grid((m + 255) / 256 > 64 ? 64 : (m + 255) / 256, n < 1024 ? n : 1024);
  matrix_scale_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(C, ldc, n, m, denom);
  B200OT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
