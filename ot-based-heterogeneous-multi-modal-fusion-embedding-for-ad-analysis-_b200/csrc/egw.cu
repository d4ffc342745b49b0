// Entropic Gromov-Wasserstein sample couplings, one label (problem) per CTA, float64, everything on chip.
//
// Replaces get_coupling_egw_ott_fixed (MRI_PET_OT_OT_per_epoch_attn.py:129-186, MRI_PET_OT.py:68-122): per label,
// n_l <= 64 MRI and m_l <= 64 PET embeddings, max-scaled squared-Euclidean point-cloud geometries
// (pointcloud.PointCloud(scale_cost="max_cost"), :155-156), ott's GromovWasserstein(epsilon, max_iterations,
// linear_solver=Sinkhorn(max_iterations)) (:168-173).  The ott-jax 0.6.0 solver is not in the tree; its published
// algorithm is restated here and in the float64 CPU restatement used by the tests (parity unpinned):
//   C1 = D(X) / max D(X), C2 = D(Y) / max D(Y), a = 1/n, b = 1/m, T0 = a b^T, f = g = 0
//   outer iteration:  M = (C1 o C1) T1 (+) (C2 o C2) T^T1 - 2 C1 T C2          (square loss, h1(x) = x, h2(y) = 2y)
//                     (f, g) <- log-domain Sinkhorn on M, eps absolute, warm-started, g update then f update,
//                               error |P^T 1 - b|_1 every 10 iterations, threshold 1e-3, <= sk_max iterations
//                     T <- exp((f (+) g - M) / eps);  cost_k = <a, f> + <b, g>
//   stop when k >= gw_min and isclose(cost_{k-2}, cost_{k-1}, rtol = gw_thr), or at gw_max iterations.
// This is SURVEY.md section 8(f)-2: the per-label problems are independent (labels = the batch dimension) and small
// enough that C1, C2, M, T and the two small GEMMs of the cost update live in one CTA's shared memory for the whole
// solve; HBM is touched to read the embeddings and to write the coupling.
#include <math.h>

#include "common.cuh"

namespace b200ot {

constexpr int ET = 256;
constexpr int kEgwMax = 64;
constexpr int kEgwLd = kEgwMax + 1;  // odd leading dimension: column walks are bank-conflict free

struct EgwArgs {
  const float* X;  // all labels' rows, concatenated: (sum n_l) x dx
  const float* Y;  // (sum m_l) x dy
  const int* xoff;  // nprob + 1 row offsets into X
  const int* yoff;
  const long long* toff;  // nprob + 1 element offsets into T
  int nprob, dx, dy;
  double eps, gw_thr, sk_thr;
  int gw_max, gw_min, sk_max, sk_check;
  double* T;     // couplings, problem l at toff[l], row-major n_l x m_l
  int* info;     // per problem: outer iterations, outer converged, last inner converged, inner iterations in total
  double* cost;  // per problem: last cost_k
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < ET / 32; ++w) t += red[w];
  return t;
}

// squared-Euclidean distances of one point cloud, divided by their maximum
__device__ void cloud_cost(const float* P, int n, int d, double* Cm, double* red) {
  double mx = 0.0;
  for (int e = threadIdx.x; e < n * n; e += ET) {
    const int i = e / n, j = e - i * n;
    const float* pi = P + (size_t)i * d;
    const float* pj = P + (size_t)j * d;
    double s = 0.0;
    for (int c = 0; c < d; ++c) {
      const double df = (double)pi[c] - (double)pj[c];
      s = fma(df, df, s);
    }
    Cm[i * kEgwLd + j] = s;
    mx = fmax(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < ET / 32; ++w) t = fmax(t, red[w]);
  const double inv = t > 0.0 ? 1.0 / t : 1.0;
  for (int e = threadIdx.x; e < n * n; e += ET) {
    const int i = e / n, j = e - i * n;
    Cm[i * kEgwLd + j] *= inv;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ET) egw_kernel(const EgwArgs p) {
  extern __shared__ __align__(16) unsigned char egw_smem[];
  double* C1 = reinterpret_cast<double*>(egw_smem);
  double* C2 = C1 + kEgwMax * kEgwLd;
  double* M = C2 + kEgwMax * kEgwLd;
  double* T = M + kEgwMax * kEgwLd;
  double* W = T + kEgwMax * kEgwLd;
  double* f = W + kEgwMax * kEgwLd;
  double* g = f + kEgwMax;
  double* r1 = g + kEgwMax;
  double* r2 = r1 + kEgwMax;
  double* rs = r2 + kEgwMax;
  double* cs = rs + kEgwMax;
  double* red = cs + kEgwMax;  // ET / 32
  __shared__ double err_sh;
  const int tid = threadIdx.x;
  const int prob = blockIdx.x;
  const int x0 = p.xoff[prob], y0 = p.yoff[prob];
  const int n = p.xoff[prob + 1] - x0, m = p.yoff[prob + 1] - y0;
  const double eps = p.eps, ieps = 1.0 / p.eps;
  const double la = -log((double)n), lb = -log((double)m);  // log a_i, log b_j (uniform marginals)
  const double bj = 1.0 / m;

  cloud_cost(p.X + (size_t)x0 * p.dx, n, p.dx, C1, red);
  cloud_cost(p.Y + (size_t)y0 * p.dy, m, p.dy, C2, red);
  for (int e = tid; e < n * m; e += ET) T[(e / m) * kEgwLd + (e % m)] = 1.0 / ((double)n * (double)m);
  for (int i = tid; i < kEgwMax; i += ET) f[i] = g[i] = 0.0;
  __syncthreads();

  double c_prev1 = 0.0, c_prev2 = 0.0;
  int outer = 0, inner_total = 0, inner_conv = 0, outer_conv = 0;
  while (outer < p.gw_max) {
    // ---- linearisation at T ---------------------------------------------------------------------------------
    for (int i = tid; i < n; i += ET) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += T[i * kEgwLd + j];
      rs[i] = s;
    }
    for (int j = tid; j < m; j += ET) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += T[i * kEgwLd + j];
      cs[j] = s;
    }
    __syncthreads();
    for (int i = tid; i < n; i += ET) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(C1[i * kEgwLd + k] * C1[i * kEgwLd + k], rs[k], s);
      r1[i] = s;
    }
    for (int j = tid; j < m; j += ET) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s = fma(C2[j * kEgwLd + k] * C2[j * kEgwLd + k], cs[k], s);
      r2[j] = s;
    }
    for (int e = tid; e < n * m; e += ET) {  // W = T C2
      const int i = e / m, j = e - i * m;
      double s = 0.0;
      for (int k = 0; k < m; ++k) s = fma(T[i * kEgwLd + k], C2[k * kEgwLd + j], s);
      W[i * kEgwLd + j] = s;
    }
    __syncthreads();
    for (int e = tid; e < n * m; e += ET) {  // M = r1 (+) r2 - 2 C1 W
      const int i = e / m, j = e - i * m;
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(C1[i * kEgwLd + k], W[k * kEgwLd + j], s);
      M[i * kEgwLd + j] = r1[i] + r2[j] - 2.0 * s;
    }
    __syncthreads();

    // ---- inner log-domain Sinkhorn on M, warm-started --------------------------------------------------------
    int it = 0, conv = 0;
    while (it < p.sk_max) {
      {  // g_j = eps log b_j - eps LSE_i((f_i - M_ij) / eps): 4 threads per column
        const int j = tid >> 2, q = tid & 3;
        double mx = -INFINITY;
        if (j < m)
          for (int i = q; i < n; i += 4) mx = fmax(mx, (f[i] - M[i * kEgwLd + j]) * ieps);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double s = 0.0;
        if (j < m)
          for (int i = q; i < n; i += 4) s += exp((f[i] - M[i * kEgwLd + j]) * ieps - mx);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        __syncthreads();  // every reader of g (previous f update) is done
        if (j < m && q == 0) g[j] = eps * lb - eps * (mx + log(s));
      }
      __syncthreads();
      {  // f_i = eps log a_i - eps LSE_j((g_j - M_ij) / eps): 4 threads per row
        const int i = tid >> 2, q = tid & 3;
        double mx = -INFINITY;
        if (i < n)
          for (int j = q; j < m; j += 4) mx = fmax(mx, (g[j] - M[i * kEgwLd + j]) * ieps);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double s = 0.0;
        if (i < n)
          for (int j = q; j < m; j += 4) s += exp((g[j] - M[i * kEgwLd + j]) * ieps - mx);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        __syncthreads();
        if (i < n && q == 0) f[i] = eps * la - eps * (mx + log(s));
      }
      __syncthreads();
      ++it;
      if (it % p.sk_check == 0) {  // |P^T 1 - b|_1
        const int j = tid >> 2, q = tid & 3;
        double s = 0.0;
        if (j < m)
          for (int i = q; i < n; i += 4) s += exp((f[i] + g[j] - M[i * kEgwLd + j]) * ieps);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const double e = (j < m && q == 0) ? fabs(s - bj) : 0.0;
        const double tot = block_sum(e, red);
        if (tid == 0) err_sh = tot;
        __syncthreads();
        if (err_sh < p.sk_thr) {
          conv = 1;
          break;
        }
      }
    }
    inner_total += it;
    inner_conv = conv;

    // ---- new coupling and its cost ---------------------------------------------------------------------------
    for (int e = tid; e < n * m; e += ET) {
      const int i = e / m, j = e - i * m;
      T[i * kEgwLd + j] = exp((f[i] + g[j] - M[i * kEgwLd + j]) * ieps);
    }
    double cpart = 0.0;
    if (tid < n) cpart += f[tid] / n;
    if (tid < m) cpart += g[tid] / m;
    const double cost = block_sum(cpart, red);
    __syncthreads();
    ++outer;
    // ott: converged(iteration) = iteration >= 2 and isclose(costs[iteration - 2], costs[iteration - 1], rtol)
    c_prev2 = c_prev1;
    c_prev1 = cost;
    if (outer >= 2) {
      const bool close = fabs(c_prev2 - c_prev1) <= 1e-8 + p.gw_thr * fabs(c_prev1);
      outer_conv = close ? 1 : 0;
      if (close && outer >= p.gw_min) break;
    }
  }

  double* Tout = p.T + p.toff[prob];
  for (int e = tid; e < n * m; e += ET) Tout[e] = T[(e / m) * kEgwLd + (e % m)];
  if (tid == 0) {
    p.info[prob * 4 + 0] = outer;
    p.info[prob * 4 + 1] = outer_conv;
    p.info[prob * 4 + 2] = inner_conv;
    p.info[prob * 4 + 3] = inner_total;
    p.cost[prob] = c_prev1;
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_egw_batched(const float* X, const float* Y, const int* xoff, const int* yoff, const long long* toff,
                       int nprob, int max_n, int max_m, int dx, int dy, float eps, int gw_max_iter, int gw_min_iter,
                       float gw_threshold, int sk_max_iter, int sk_check_every, float sk_threshold, double* T,
                       int* info4, double* cost, void* stream) {
  if (!X || !Y || !xoff || !yoff || !toff || !T || !info4 || !cost || nprob <= 0 || dx <= 0 || dy <= 0)
    return B200OT_E_INVALID;
  if (!(eps > 0.f) || gw_max_iter < 1 || sk_max_iter < 1 || sk_check_every < 1) return B200OT_E_INVALID;
  if (max_n < 1 || max_m < 1 || max_n > kEgwMax || max_m > kEgwMax) return B200OT_E_UNSUPPORTED;
  EgwArgs p;
  p.X = X;
  p.Y = Y;
  p.xoff = xoff;
  p.yoff = yoff;
  p.toff = toff;
  p.nprob = nprob;
  p.dx = dx;
  p.dy = dy;
  p.eps = (double)eps;
  p.gw_thr = (double)gw_threshold;
  p.sk_thr = (double)sk_threshold;
  p.gw_max = gw_max_iter;
  p.gw_min = gw_min_iter;
  p.sk_max = sk_max_iter;
  p.sk_check = sk_check_every;
  p.T = T;
  p.info = info4;
  p.cost = cost;
  const size_t smem = ((size_t)5 * kEgwMax * kEgwLd + 6 * kEgwMax + ET / 32) * sizeof(double);
  static PerDeviceOnce attr_once;  // function attributes are per device
  if (attr_once.first()) {
    B200OT_CUDA_OK(cudaFuncSetAttribute(egw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  egw_kernel<<<nprob, ET, smem, static_cast<cudaStream_t>(stream)>>>(p);
  B200OT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
