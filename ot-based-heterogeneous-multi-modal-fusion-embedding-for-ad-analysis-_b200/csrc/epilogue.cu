// Epilogues of the OT solve: plan materialisation, transport cost, plan application
// (barycentric projection / `pet @ T.t()`), cosine fusion loss.
//
//   plan        diag(u) K diag(v) of perturbot/perturbot/match/utils.py:111-115 and ott's
//               `.matrix` (perturbot/perturbot/match/fot.py:134): P_ij = exp((f_i+g_j-C_ij)/eps)
//   ot_cost     sum(M * Tv) of perturbot/perturbot/match/fot.py:137
//   apply_plan  (T / rowsum) @ Y of perturbot/perturbot/eval/match.py:202-206 (normalise=1) and
//               pet @ T.t() of MRI_PET_OT_OT_per_epoch_attn.py:728 (transposed form, raw plan)
//   cosine_loss 1 - mean_i cos(x_i, y_i) of MRI_PET_OT_nojax.py:552-560
// None of them materialises P unless asked to (b200ot_plan).
#include <math.h>

#include "common.cuh"

namespace b200ot {

__global__ void __launch_bounds__(256) plan_kernel(const float* __restrict__ C, long long ldc, int n,
                                                   int m, const float* __restrict__ f,
                                                   const float* __restrict__ g, float k, float* P,
                                                   long long ldp) {
  for (long long r = blockIdx.y; r < n; r += gridDim.y) {
    const float fi = f[r] * k;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < m; j += gridDim.x * 256)
      P[r * ldp + j] = exp2f(fmaf(C[r * ldc + j], -k, fmaf(g[j], k, fi)));
  }
}

// <P, C> without materialising P.  Every block leaves ONE double (its rows in fixed order, fixed reduction tree)
// in out[1 + block]; the last block to finish (ticket in out[1 + kOtCostBlocks]) adds the partials in block order,
// so the value is bit-reproducible run to run -- no floating-point atomics.
constexpr int kOtCostBlocks = 148 * 8;
__global__ void __launch_bounds__(256) ot_cost_kernel(const float* __restrict__ C, long long ldc, int n,
                                                      int m, const float* __restrict__ f,
                                                      const float* __restrict__ g, float k,
                                                      double* out) {
  double acc = 0.0;
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const float fi = f[r] * k;
    float part = 0.f;
    for (int j = threadIdx.x; j < m; j += 256) {
      const float c = C[r * ldc + j];
      part = fmaf(exp2f(fmaf(c, -k, fmaf(g[j], k, fi))), c, part);
    }
    acc += (double)part;
  }
  acc = warp_sum(acc);
  __shared__ double sh[8];
  __shared__ int is_last;
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[1 + blockIdx.x] = t;
    __threadfence();
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(out + 1 + kOtCostBlocks);
    is_last = (atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double part = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) part += ((volatile double*)out)[1 + i];
  part = warp_sum(part);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[0] = t;
  }
}

// The per-step plan guard of the reference (MRI_PET_OT_nojax.py:704-715): NaN -> 1e-8, then every row divided by
// its sum with zero sums replaced by 1e-8.  The plan comes either from potentials (T == nullptr: entries are
// re-evaluated, never stored unnormalised) or from a materialised matrix T.  One block per row, two passes.
__global__ void __launch_bounds__(256) plan_guard_rownorm_kernel(const float* __restrict__ C, long long ldc, int n,
                                                                 int m, const float* __restrict__ f,
                                                                 const float* __restrict__ g, float k,
                                                                 const float* __restrict__ T, long long ldt,
                                                                 float* __restrict__ P, long long ldp) {
  __shared__ float sh[8];
  __shared__ float s_inv;
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const float fi = T ? 0.f : f[r] * k;
    auto entry = [&](int j) {
      float v = T ? T[r * ldt + j] : exp2f(fmaf(C[r * ldc + j], -k, fmaf(g[j], k, fi)));
      return (v != v) ? 1e-8f : v;
    };
    float part = 0.f;
    for (int j = threadIdx.x; j < m; j += 256) part += entry(j);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += sh[w];
      s_inv = 1.f / (t == 0.f ? 1e-8f : t);
    }
    __syncthreads();
    const float inv = s_inv;
    for (int j = threadIdx.x; j < m; j += 256) P[r * ldp + j] = entry(j) * inv;
    __syncthreads();
  }
}

// Z(i, c) = [1/rowsum_i] sum_j P(i,j) V(j,c), P(i,j) = 2^(fs_i + gs_j - k C(i,j)); C is addressed
// with explicit strides so the same kernel serves P V (sci = ldc, scj = 1) and P^T U
// (sci = 1, scj = ldc with the roles of f and g swapped).
constexpr int AT_I = 64, AT_C = 64, AT_J = 16;
__global__ void __launch_bounds__(256) apply_plan_kernel(const float* __restrict__ C, long long sci,
                                                         long long scj, int n, int m,
                                                         const float* __restrict__ f,
                                                         const float* __restrict__ g, float k,
                                                         const float* __restrict__ V, long long ldv,
                                                         int dv, int normalise, float* Z,
                                                         long long ldz) {
  __shared__ float Ps[AT_J][AT_I + 1];
  __shared__ float Vs[AT_J][AT_C + 1];
  __shared__ float rs[AT_I];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.y * AT_I, c0 = blockIdx.x * AT_C;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each (strided by 16)
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  float rsum = 0.f;  // threads 0..63 own one row's plan sum
  const bool jfast = (scj == 1);
  for (int j0 = 0; j0 < m; j0 += AT_J) {
#pragma unroll
    for (int e = 0; e < (AT_I * AT_J) / 256; ++e) {
      const int idx = e * 256 + tid;
      int ii, jj;
      if (jfast) {
        jj = idx & (AT_J - 1);
        ii = idx >> 4;
      } else {
        ii = idx & (AT_I - 1);
        jj = idx >> 6;
      }
      const int gi = i0 + ii, gj = j0 + jj;
      float p = 0.f;
      if (gi < n && gj < m) p = exp2f(fmaf(C[gi * sci + gj * scj], -k, fmaf(f[gi], k, g[gj] * k)));
      Ps[jj][ii] = p;
    }
#pragma unroll
    for (int e = 0; e < (AT_J * AT_C) / 256; ++e) {
      const int idx = e * 256 + tid;
      const int cc = idx & (AT_C - 1), jj = idx >> 6;
      const int gj = j0 + jj, gc = c0 + cc;
      Vs[jj][cc] = (gj < m && gc < dv) ? V[(long long)gj * ldv + gc] : 0.f;
    }
    __syncthreads();
    if (tid < AT_I) {
#pragma unroll
      for (int jj = 0; jj < AT_J; ++jj) rsum += Ps[jj][tid];
    }
#pragma unroll
    for (int jj = 0; jj < AT_J; ++jj) {
      float pv[4], vv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pv[r] = Ps[jj][ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) vv[c] = Vs[jj][tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(pv[r], vv[c], acc[r][c]);
    }
    __syncthreads();
  }
  if (tid < AT_I) rs[tid] = rsum;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ii = ty + 16 * r, gi = i0 + ii;
    if (gi >= n) continue;
    float sc = 1.f;
    if (normalise) {
      const float s = rs[ii];
      sc = 1.f / (s == 0.f ? 1e-30f : s);  // marg == 0 -> 1e-30 (eval/match.py:203-204)
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int gc = c0 + tx + 16 * c;
      if (gc < dv) Z[(long long)gi * ldz + gc] = acc[r][c] * sc;
    }
  }
}

__global__ void __launch_bounds__(256) cosine_loss_kernel(const float* __restrict__ A, long long lda,
                                                          const float* __restrict__ B, long long ldb,
                                                          int rows, int d, float* out) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float contrib = 0.f;
  if (warp < rows) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float x = A[warp * lda + c], y = B[warp * ldb + c];
      ab = fmaf(x, y, ab);
      aa = fmaf(x, x, aa);
      bb = fmaf(y, y, bb);
    }
    ab = warp_sum(ab);
    aa = warp_sum(aa);
    bb = warp_sum(bb);
    // F.normalize(eps=1e-12) then cosine_similarity(eps=1e-8)
    const float na = fmaxf(sqrtf(aa), 1e-12f), nb = fmaxf(sqrtf(bb), 1e-12f);
    const float num = ab / (na * nb);
    const float den = fmaxf((sqrtf(aa) / na) * (sqrtf(bb) / nb), 1e-8f);
    contrib = num / den;
  }
  __shared__ float sh[8];
  if (lane == 0) sh[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(out, -t / (float)rows);
  }
}
__global__ void set_one_kernel(float* p) { *p = 1.f; }

// FOSCTTM (perturbot/perturbot/eval/utils.py:18-45): for sample i, the rank of its true match among all
// distances d(pred_i, true_j), ties at their mean position, divided by n - 1.  D is the n x n matrix of
// squared distances (monotone in the reference's Euclidean distances).  One warp per row.
__global__ void __launch_bounds__(256) foscttm_kernel(const float* __restrict__ D, long long ldd, int n,
                                                      float* __restrict__ out) {
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* d = D + (long long)row * ldd;
  const float self = d[row];
  int less = 0, equal = 0;
  for (int j = lane; j < n; j += 32) {
    const float v = d[j];
    less += v < self;
    equal += v == self;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    less += __shfl_xor_sync(0xffffffffu, less, o);
    equal += __shfl_xor_sync(0xffffffffu, equal, o);
  }
  if (lane == 0) out[row] = n > 1 ? ((float)less + 0.5f * (float)(equal - 1)) / (float)(n - 1) : 0.f;
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_plan(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                float* P, int ldp, void* stream) {
  if (!C || !f || !g || !P || n <= 0 || m <= 0 || ldc < m || ldp < m || !(eps > 0.f))
    return B200OT_E_INVALID;
  dim3 grid((m + 255) / 256 > 32 ? 32 : (m + 255) / 256, n < 2048 ? n : 2048);
  plan_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(C, ldc, n, m, f, g, kLog2e / eps, P, ldp);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_ot_cost(const float* C, int ldc, int n, int m, const float* f, const float* g,
                   float eps, double* out, void* stream) {
  if (!C || !f || !g || !out || n <= 0 || m <= 0 || ldc < m || !(eps > 0.f)) return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200OT_CUDA_OK(cudaMemsetAsync(out, 0, B200OT_OT_COST_DOUBLES * sizeof(double), s));
  const int grid = n < kOtCostBlocks ? n : kOtCostBlocks;
  ot_cost_kernel<<<grid, 256, 0, s>>>(C, ldc, n, m, f, g, kLog2e / eps, out);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_plan_guard_rownorm(const float* C, int ldc, int n, int m, const float* f, const float* g, float eps,
                              const float* T, int ldt, float* P, int ldp, void* stream) {
  if (!P || n <= 0 || m <= 0 || ldp < m) return B200OT_E_INVALID;
  if (T) {
    if (ldt < m) return B200OT_E_INVALID;
  } else if (!C || !f || !g || ldc < m || !(eps > 0.f)) {
    return B200OT_E_INVALID;
  }
  const int grid = n < 148 * 8 ? n : 148 * 8;
  plan_guard_rownorm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      C, ldc, n, m, f, g, T ? 0.f : kLog2e / eps, T, ldt, P, ldp);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_apply_plan(const float* C, int ldc, int n, int m, const float* f, const float* g,
                      float eps, const float* V, int ldv, int dv, int normalise, float* Z,
                      int ldz, void* stream) {
  if (!C || !f || !g || !V || !Z || n <= 0 || m <= 0 || dv <= 0 || ldc < m || ldv < dv || ldz < dv ||
      !(eps > 0.f))
    return B200OT_E_INVALID;
  dim3 grid((dv + AT_C - 1) / AT_C, (n + AT_I - 1) / AT_I);
  apply_plan_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      C, ldc, 1, n, m, f, g, kLog2e / eps, V, ldv, dv, normalise, Z, ldz);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_apply_plan_t(const float* C, int ldc, int n, int m, const float* f, const float* g,
                        float eps, const float* U, int ldu, int du, int normalise, float* Z, int ldz,
                        void* stream) {
  if (!C || !f || !g || !U || !Z || n <= 0 || m <= 0 || du <= 0 || ldc < m || ldu < du || ldz < du ||
      !(eps > 0.f))
    return B200OT_E_INVALID;
  // Z (m x du) = P^T U: rows of the product are columns of C
  dim3 grid((du + AT_C - 1) / AT_C, (m + AT_I - 1) / AT_I);
  apply_plan_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      C, 1, ldc, m, n, g, f, kLog2e / eps, U, ldu, du, normalise, Z, ldz);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_foscttm(const float* D, int ldd, int n, float* out, void* stream) {
  if (!D || !out || n <= 0 || ldd < n) return B200OT_E_INVALID;
  foscttm_kernel<<<(n * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(D, ldd, n, out);
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_cosine_loss(const float* A, int lda, const float* B, int ldb, int rows, int d,
                       float* out, void* stream) {
  if (!A || !B || !out || rows <= 0 || d <= 0 || lda < d || ldb < d) return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  set_one_kernel<<<1, 1, 0, s>>>(out);
  B200OT_LAUNCH_OK();
  cosine_loss_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(A, lda, B, ldb, rows, d, out);
  B200OT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
