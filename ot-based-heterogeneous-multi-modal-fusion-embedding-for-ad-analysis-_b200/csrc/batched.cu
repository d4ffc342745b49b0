// Batched small-problem Sinkhorn: one problem per CTA, float64, kernel domain.
//
// This is the per-training-step minibatch OT of the reference
// (MRI_PET_OT_nojax.py:679-715 -> get_feature_coupling_pot :91-145 -> ot.sinkhorn :143) run for a
// whole batch of independent problems at once.  The arithmetic is POT's sinkhorn_knopp as
// mirrored in perturbot/perturbot/match/utils.py:6-115, in float64 like the reference:
//   K = exp(-M/reg);  u = 1/n, v = 1/m;  Kp = (1/a) K
//   loop:  v = b / (K^T u);  u = 1 / (Kp v)
//          on K^T u == 0 or NaN/Inf in u, v: restore the previous u, v and stop  (:55-79)
//          every check_every-th iteration: err = |colsum(diag(u) K diag(v)) - b|  (:80-89)
//   P = diag(u) K diag(v)                                                          (:111-115)
// K stays in shared memory for the whole solve (n, m <= 128), so a solve touches HBM only to
// read the embeddings / cost and to write the plan.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace b200ot {

constexpr int BT = 256;  // threads per problem

struct BatchedArgs {
  const float* C;  // batch x n x m, or null
  const float* X;  // batch x n x d
  const float* Y;  // batch x m x d
  int batch, n, m, d;
  const float* a;  // n (shared by all problems)
  const float* b;  // m
  double reg;
  int max_iter, check_every, check_phase, err_norm, stop_inclusive;
  double tol;
  float* P;      // batch x n x m
  double* u;     // batch x n  (may be null)
  double* v;     // batch x m  (may be null)
  int* n_iter;   // batch      (may be null)
  float* err;    // batch      (may be null)
};

__device__ __forceinline__ bool bad_value(double x) { return isnan(x) || isinf(x); }

__global__ void __launch_bounds__(BT) sinkhorn_batched_kernel(const BatchedArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m;
  const int ldk = m | 1;  // odd row stride: the row pass reads K down a column of banks
  double* K = reinterpret_cast<double*>(smem_raw);      // n x ldk
  double* u = K + (size_t)n * ldk;                      // n
  double* v = u + n;                                    // m
  double* up = v + m;                                   // previous iterates
  double* vp = up + n;
  double* ktu = vp + m;                                 // m
  double* scratch = ktu + m;                            // BT doubles
  __shared__ int flag_sh;
  __shared__ double err_sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BT / 32;

  for (int prob = blockIdx.x; prob < p.batch; prob += gridDim.x) {
    // ---- K = exp(-M/reg) ----
    if (p.C) {
      const float* Cp = p.C + (size_t)prob * n * m;
      for (int e = tid; e < n * m; e += BT) {
        const int i = e / m, j = e - i * m;
        K[i * ldk + j] = exp(-(double)Cp[e] / p.reg);
      }
    } else if (n <= 64 && m <= 64) {
      // squared-Euclidean cost from the embeddings in float64, tiled through shared memory:
      // 16 x 16 threads, each owning a 4 x 4 block (rows ty + 16 r, columns tx + 16 c) of the 64 x 64 cost
      const float* Xp = p.X + (size_t)prob * n * p.d;
      const float* Yp = p.Y + (size_t)prob * m * p.d;
      float* Xs = reinterpret_cast<float*>(scratch + BT);  // [64][33]
      float* Ys = Xs + 64 * 33;
      const int tx = tid & 15, ty = tid >> 4;
      double acc[4][4], xx[4], yy[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        xx[r] = yy[r] = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
      }
      for (int kc = 0; kc < p.d; kc += 32) {
        for (int e = tid; e < 64 * 32; e += BT) {
          const int i = e >> 5, kk = e & 31;
          const bool kin = kc + kk < p.d;
          Xs[i * 33 + kk] = (i < n && kin) ? Xp[(size_t)i * p.d + kc + kk] : 0.f;
          Ys[i * 33 + kk] = (i < m && kin) ? Yp[(size_t)i * p.d + kc + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
          double xv[4], yv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            xv[r] = (double)Xs[(ty + 16 * r) * 33 + kk];
            yv[r] = (double)Ys[(tx + 16 * r) * 33 + kk];
            xx[r] = fma(xv[r], xv[r], xx[r]);
            yy[r] = fma(yv[r], yv[r], yy[r]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fma(xv[r], yv[c], acc[r][c]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = ty + 16 * r, j = tx + 16 * c;
          if (i < n && j < m) K[i * ldk + j] = exp(-((xx[r] + yy[c]) - 2.0 * acc[r][c]) / p.reg);
        }
    } else {
      // squared-Euclidean cost from the embeddings, accumulated in float64 (generic sizes)
      const float* Xp = p.X + (size_t)prob * n * p.d;
      const float* Yp = p.Y + (size_t)prob * m * p.d;
      for (int e = tid; e < n * m; e += BT) {
        const int i = e / m, j = e - i * m;
        const float* x = Xp + (size_t)i * p.d;
        const float* y = Yp + (size_t)j * p.d;
        double xx = 0.0, yy = 0.0, xy = 0.0;
        for (int c = 0; c < p.d; ++c) {
          const double xv = x[c], yv = y[c];
          xx = fma(xv, xv, xx);
          yy = fma(yv, yv, yy);
          xy = fma(xv, yv, xy);
        }
        K[i * ldk + j] = exp(-((xx + yy) - 2.0 * xy) / p.reg);
      }
    }
    for (int i = tid; i < n; i += BT) u[i] = 1.0 / n;
    for (int j = tid; j < m; j += BT) v[j] = 1.0 / m;
    if (tid == 0) {
      flag_sh = 0;
      err_sh = 1.0;
    }
    __syncthreads();

    // Both mat-vecs use the same layout: thread (index, part) sums a strided slice (PARTS = BT / 64 slices for up
    // to 64 indices, 2 for up to 128), the slices are folded through shared memory by the first threads, which also
    // keep the previous iterate and apply POT's numerical guard.  Four block barriers per iteration.  (The first
    // version ran the row pass one warp per row: eight serial rows per warp, each with a double-precision shuffle
    // reduction and a division -- 5k cycles per iteration where the arithmetic needs ~300.)
    const int cparts = BT / m > 0 ? BT / m : 1;
    const int rparts = BT / n > 0 ? BT / n : 1;
    int cpt = 0;
    double err = 1.0;
    while (cpt < p.max_iter) {
      // ---- v = b / (K^T u) ----
      {
        const int j = tid % m, part = tid / m;
        double s0 = 0.0, s1 = 0.0;  // two chains: the loads of one overlap the FMA latency of the other
        if (part < cparts) {
          const double* kp = K + part * ldk + j;
          const int step = cparts * ldk;
          int i = part;
          for (; i + cparts < n; i += 2 * cparts, kp += 2 * step) {
            s0 = fma(kp[0], u[i], s0);
            s1 = fma(kp[step], u[i + cparts], s1);
          }
          if (i < n) s0 = fma(kp[0], u[i], s0);
        }
        scratch[tid] = s0 + s1;
        __syncthreads();
        if (tid < m) {
          double t = 0.0;
          for (int q = 0; q < cparts; ++q) t += scratch[q * m + tid];
          vp[tid] = v[tid];
          ktu[tid] = t;
          const double vn = (double)p.b[tid] / t;
          v[tid] = vn;
          if (t == 0.0 || bad_value(vn)) flag_sh = 1;
        }
        __syncthreads();
      }
      // ---- u = 1 / (Kp v), Kp = (1/a) K ----
      {
        const int i = tid % n, part = tid / n;
        double s0 = 0.0, s1 = 0.0;
        if (part < rparts) {
          const double* kp = K + i * ldk;
          int j = part;
          for (; j + rparts < m; j += 2 * rparts) {
            s0 = fma(kp[j], v[j], s0);
            s1 = fma(kp[j + rparts], v[j + rparts], s1);
          }
          if (j < m) s0 = fma(kp[j], v[j], s0);
        }
        scratch[tid] = s0 + s1;
        __syncthreads();
        if (tid < n) {
          double t = 0.0;
          for (int q = 0; q < rparts; ++q) t += scratch[q * n + tid];
          up[tid] = u[tid];
          const double un = (double)p.a[tid] / t;  // = 1 / (Kp v)_i
          u[tid] = un;
          if (bad_value(un)) flag_sh = 1;
        }
        __syncthreads();
      }
      // numerical guard (utils.py:55-79): restore the previous iterates and stop
      if (flag_sh) {
        for (int i = tid; i < n; i += BT) u[i] = up[i];
        for (int j = tid; j < m; j += BT) v[j] = vp[j];
        __syncthreads();
        break;
      }
      if (cpt % p.check_every == ((p.check_phase + p.check_every - 1) % p.check_every)) {
        // column marginal of diag(u) K diag(v)
        const int j = tid % m, part = tid / m;
        double s = 0.0;
        if (part < cparts)
          for (int i = part; i < n; i += cparts) s = fma(u[i], K[i * ldk + j] * v[j], s);
        scratch[tid] = s;
        __syncthreads();
        double e = 0.0;
        if (tid < m) {
          double t = 0.0;
          for (int q = 0; q < cparts; ++q) t += scratch[q * m + tid];
          const double dlt = t - (double)p.b[tid];
          e = p.err_norm == B200OT_NORM_L1 ? fabs(dlt) : dlt * dlt;
        }
        e = warp_sum(e);
        __syncthreads();
        if (lane == 0) scratch[warp] = e;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w = 0; w < NW; ++w) t += scratch[w];
          err_sh = p.err_norm == B200OT_NORM_L2 ? sqrt(t) : t;
        }
        __syncthreads();
        err = err_sh;
        const bool stop = p.stop_inclusive ? (err <= p.tol) : (err < p.tol);
        if (stop) {
          ++cpt;
          break;
        }
      }
      ++cpt;
    }

    // ---- outputs ----
    float* Pp = p.P + (size_t)prob * n * m;
    for (int e = tid; e < n * m; e += BT) {
      const int i = e / m, j = e - i * m;
      Pp[e] = (float)(u[i] * K[i * ldk + j] * v[j]);
    }
    if (p.u)
      for (int i = tid; i < n; i += BT) p.u[(size_t)prob * n + i] = u[i];
    if (p.v)
      for (int j = tid; j < m; j += BT) p.v[(size_t)prob * m + j] = v[j];
    if (tid == 0) {
      if (p.n_iter) p.n_iter[prob] = cpt;
      if (p.err) p.err[prob] = (float)err;
    }
    __syncthreads();
  }
}

// =============================================================================
// n, m <= 64: the Gibbs kernel lives in REGISTERS
// =============================================================================
// The general kernel above re-reads K (32 KiB of float64 at 64 x 64) from shared memory twice per iteration and is
// bound by those wavefronts (measured 1.85 k cycles per problem-iteration).  Here thread (ty, tx) of a 16 x 16 grid
// owns the 4 x 4 strided block K[ty + 16 r][tx + 16 c] in registers for the whole solve -- the block its
// float64 accumulators already hold when the cost is built from the embeddings -- and only the vectors travel:
//   column pass  16 DFMA, fold the two ty of a warp with ONE exchange step (each lane gives two columns away and
//                keeps two), 2 x 8-byte stores per lane, 64 threads fold the 8 warps' partials and update v;
//   row pass     16 DFMA, transposing butterfly over the 16 tx of a half-warp (5 exchanges for 4 rows instead of
//                16), the lanes that end up owning a row update u -- 64 threads, one division each.
// Same arithmetic as POT's loop (utils.py:45-89) in float64; only the order of the additions differs.
__device__ __forceinline__ double shfl_xor_d(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

__global__ void __launch_bounds__(BT, 2) sinkhorn_batched64_kernel(const BatchedArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m;
  double* u = reinterpret_cast<double*>(smem_raw);  // 64
  double* v = u + 64;
  double* up = v + 64;
  double* vp = up + 64;
  double* part = vp + 64;          // [8 warps][64 columns]
  double* red = part + 8 * 64;     // 8
  float* Xs = reinterpret_cast<float*>(red + 8);  // [64][33] (embedding path only)
  float* Ys = Xs + 64 * 33;
  __shared__ int flag_sh;
  __shared__ double err_sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const bool hi = (lane & 16) != 0;  // the odd ty of this warp

  for (int prob = blockIdx.x; prob < p.batch; prob += gridDim.x) {
    double K[4][4];
    // ---- K = exp(-M/reg), straight into registers ----
    if (p.C) {
      const float* Cp = p.C + (size_t)prob * n * m;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = ty + 16 * r, j = tx + 16 * c;
          K[r][c] = (i < n && j < m) ? exp(-(double)Cp[(size_t)i * m + j] / p.reg) : 0.0;
        }
    } else {
      const float* Xp = p.X + (size_t)prob * n * p.d;
      const float* Yp = p.Y + (size_t)prob * m * p.d;
      double acc[4][4], xx[4], yy[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        xx[r] = yy[r] = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
      }
      for (int kc = 0; kc < p.d; kc += 32) {
        for (int e = tid; e < 64 * 32; e += BT) {
          const int i = e >> 5, kk = e & 31;
          const bool kin = kc + kk < p.d;
          Xs[i * 33 + kk] = (i < n && kin) ? Xp[(size_t)i * p.d + kc + kk] : 0.f;
          Ys[i * 33 + kk] = (i < m && kin) ? Yp[(size_t)i * p.d + kc + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
          double xv[4], yv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            xv[r] = (double)Xs[(ty + 16 * r) * 33 + kk];
            yv[r] = (double)Ys[(tx + 16 * r) * 33 + kk];
            xx[r] = fma(xv[r], xv[r], xx[r]);
            yy[r] = fma(yv[r], yv[r], yy[r]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fma(xv[r], yv[c], acc[r][c]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = ty + 16 * r, j = tx + 16 * c;
          K[r][c] = (i < n && j < m) ? exp(-((xx[r] + yy[c]) - 2.0 * acc[r][c]) / p.reg) : 0.0;
        }
    }
    if (tid < 64) {
      u[tid] = 1.0 / n;
      v[tid] = 1.0 / m;
    }
    if (tid == 0) {
      flag_sh = 0;
      err_sh = 1.0;
    }
    __syncthreads();

    // column sums of diag(w) K for a row vector w in shared memory: afterwards part[8][64] holds the warps' partials
    auto column_partials = [&](const double* w) {
      double wr[4], pc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wr[r] = w[ty + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double sacc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) sacc = fma(K[r][c], wr[r], sacc);
        pc[c] = sacc;
      }
      // fold the two ty of the warp: the even ty keeps columns c = 0, 1, the odd ty keeps c = 2, 3
      const double g0 = shfl_xor_d(hi ? pc[0] : pc[2], 16);
      const double g1 = shfl_xor_d(hi ? pc[1] : pc[3], 16);
      const double k0 = (hi ? pc[2] : pc[0]) + g0;
      const double k1 = (hi ? pc[3] : pc[1]) + g1;
      const int cbase = hi ? 2 : 0;
      part[warp * 64 + tx + 16 * cbase] = k0;
      part[warp * 64 + tx + 16 * (cbase + 1)] = k1;
    };

    int cpt = 0;
    double err = 1.0;
    while (cpt < p.max_iter) {
      // ---- v = b / (K^T u) ----
      column_partials(u);
      __syncthreads();
      if (tid < 64) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w * 64 + tid];
        if (tid < m) {
          vp[tid] = v[tid];
          const double vn = (double)p.b[tid] / t;
          v[tid] = vn;
          if (t == 0.0 || bad_value(vn)) flag_sh = 1;
        }
      }
      __syncthreads();
      // ---- u = 1 / (Kp v), Kp = (1/a) K ----
      {
        double vr[4], pr[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) vr[c] = v[tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          double sacc = 0.0;
#pragma unroll
          for (int c = 0; c < 4; ++c) sacc = fma(K[r][c], vr[c], sacc);
          pr[r] = sacc;
        }
        // transposing butterfly over the 16 tx (lane bits 0..3): bit 3 picks rows {0,1} / {2,3}, bit 2 picks one
        const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        const double g0 = shfl_xor_d(b3 ? pr[0] : pr[2], 8);
        const double g1 = shfl_xor_d(b3 ? pr[1] : pr[3], 8);
        const double q0 = (b3 ? pr[2] : pr[0]) + g0;
        const double q1 = (b3 ? pr[3] : pr[1]) + g1;
        const double g2 = shfl_xor_d(b2 ? q0 : q1, 4);
        double tot = (b2 ? q1 : q0) + g2;
        tot += shfl_xor_d(tot, 2);
        tot += shfl_xor_d(tot, 1);
        if ((lane & 3) == 0) {
          const int i = ty + 16 * ((b3 ? 2 : 0) + (b2 ? 1 : 0));
          if (i < n) {
            up[i] = u[i];
            const double un = (double)p.a[i] / tot;  // = 1 / (Kp v)_i
            u[i] = un;
            if (bad_value(un)) flag_sh = 1;
          }
        }
      }
      __syncthreads();
      // numerical guard (utils.py:55-79): restore the previous iterates and stop
      if (flag_sh) {
        if (tid < 64) {
          if (tid < n) u[tid] = up[tid];
          if (tid < m) v[tid] = vp[tid];
        }
        __syncthreads();
        break;
      }
      if (cpt % p.check_every == ((p.check_phase + p.check_every - 1) % p.check_every)) {
        // column marginal of diag(u) K diag(v) with the NEW u
        column_partials(u);
        __syncthreads();
        double e = 0.0;
        if (tid < m) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) t += part[w * 64 + tid];
          const double dlt = t * v[tid] - (double)p.b[tid];
          e = p.err_norm == B200OT_NORM_L1 ? fabs(dlt) : dlt * dlt;
        }
        e = warp_sum(e);
        if (lane == 0) red[warp] = e;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w = 0; w < 8; ++w) t += red[w];
          err_sh = p.err_norm == B200OT_NORM_L2 ? sqrt(t) : t;
        }
        __syncthreads();
        err = err_sh;
        const bool stop = p.stop_inclusive ? (err <= p.tol) : (err < p.tol);
        if (stop) {
          ++cpt;
          break;
        }
      }
      ++cpt;
    }

    // ---- outputs ----
    float* Pp = p.P + (size_t)prob * n * m;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = ty + 16 * r, j = tx + 16 * c;
        if (i < n && j < m) Pp[(size_t)i * m + j] = (float)(u[i] * K[r][c] * v[j]);
      }
    if (p.u && tid < n) p.u[(size_t)prob * n + tid] = u[tid];
    if (p.v && tid < m) p.v[(size_t)prob * m + tid] = v[tid];
    if (tid == 0) {
      if (p.n_iter) p.n_iter[prob] = cpt;
      if (p.err) p.err[prob] = (float)err;
    }
    __syncthreads();
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_sinkhorn_batched(const float* C, const float* X, const float* Y, int batch, int n,
                            int m, int d, const float* a, const float* b,
                            const b200ot_params* prm, float* P, double* u, double* v, int* n_iter,
                            float* err, void* stream) {
  if (!prm || !a || !b || !P || batch <= 0 || n <= 0 || m <= 0) return B200OT_E_INVALID;
  if (!C && !(X && Y && d > 0)) return B200OT_E_INVALID;
  if (n > 128 || m > 128 || m < 2) return B200OT_E_UNSUPPORTED;
  if (!(prm->eps > 0.f)) return B200OT_E_INVALID;
  BatchedArgs p;
  p.C = C;
  p.X = X;
  p.Y = Y;
  p.batch = batch;
  p.n = n;
  p.m = m;
  p.d = d;
  p.a = a;
  p.b = b;
  p.reg = (double)prm->eps;
  p.max_iter = prm->max_iter;
  p.check_every = prm->check_every > 0 ? prm->check_every : 1;
  p.check_phase = prm->check_phase;
  p.err_norm = prm->err_norm;
  p.stop_inclusive = prm->stop_inclusive;
  p.tol = (double)prm->tol;
  p.P = P;
  p.u = u;
  p.v = v;
  p.n_iter = n_iter;
  p.err = err;
  if (n <= 64 && m <= 64) {  // Gibbs kernel in registers (B200OT_BATCHED_REG=0 keeps the shared-memory form)
    static int want = -1;
    if (want < 0) {
      const char* e = getenv("B200OT_BATCHED_REG");
      want = (e && e[0] == '0') ? 0 : 1;
    }
    if (want) {
      const size_t smem64 = (4 * 64 + 8 * 64 + 8) * sizeof(double) + (C ? 0 : 2 * 64 * 33 * sizeof(float));
      int grid = sm_count() * 2;
      if (grid > batch) grid = batch;
      sinkhorn_batched64_kernel<<<grid, BT, smem64, static_cast<cudaStream_t>(stream)>>>(p);
      B200OT_LAUNCH_OK();
      return 0;
    }
  }
  const int ldk = m | 1;
  const size_t smem = ((size_t)n * ldk + 2 * (size_t)n + 3 * (size_t)m + BT) * sizeof(double) +
                      ((!C && n <= 64 && m <= 64) ? 2 * 64 * 33 * sizeof(float) : 0);
  static PerDeviceOnce attr_once;  // function attributes are per device
  if (attr_once.first()) {
    B200OT_CUDA_OK(cudaFuncSetAttribute(sinkhorn_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
  }
  if (smem > 200 * 1024) return B200OT_E_UNSUPPORTED;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  int grid = sm_count() * per_sm;
  if (grid > batch) grid = batch;
  sinkhorn_batched_kernel<<<grid, BT, smem, static_cast<cudaStream_t>(stream)>>>(p);
  B200OT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
