// Token-mixing core of the reference's attention fusion.
//
// The reference fuses [mri_feat, OT(pet), pet2mri(pet)] with a transformer encoder block whose
// sequence length is 3 (MRI_PET_OT_OT_per_epoch_attn.py:731-740, SelfAttentionBlock :523-549; the
// nojax variant uses a single token, MRI_PET_OT_nojax.py:664-669).  For S <= 4 tokens the
// softmax(QK^T/sqrt(dh))V step is a handful of dot products per (sample, head): one warp handles one
// (sample, head) pair end to end, forward and backward, instead of three batched-GEMM launches.
// The dense projections around it (in_proj, out_proj, FFN) are plain library GEMMs in the host module.
//
// Layout: qkv is (S, B, 3E) row-major as produced by nn.MultiheadAttention's in_proj with
// batch_first=False; out is (S, B, E).  probs (B, H, S, S) is saved for the backward pass.
#include <math.h>

#include "common.cuh"

namespace b200ot {

constexpr int kMaxTokens = 4;

template <int S>
__global__ void __launch_bounds__(256) token_attn_fwd_kernel(const float* __restrict__ qkv, int B, int E, int H,
                                                             const float* __restrict__ keep, float keep_scale,
                                                             float* __restrict__ out, float* __restrict__ probs) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  const int b = wid / H, h = wid - b * H;
  const int dh = E / H;
  const float scale = rsqrtf((float)dh);
  float sc[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i)
#pragma unroll
    for (int j = 0; j < S; ++j) sc[i][j] = 0.f;
  for (int d = lane; d < dh; d += 32) {
    float q[S], k[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const float* row = qkv + ((size_t)t * B + b) * 3 * E + h * dh + d;
      q[t] = row[0];
      k[t] = row[E];
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int j = 0; j < S; ++j) sc[i][j] = fmaf(q[i], k[j], sc[i][j]);
  }
  float p[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      sc[i][j] = warp_sum(sc[i][j]) * scale;
      mx = fmaxf(mx, sc[i][j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      p[i][j] = __expf(sc[i][j] - mx);
      den += p[i][j];
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
      p[i][j] /= den;
      if (lane == 0) probs[(((size_t)b * H + h) * S + i) * S + j] = p[i][j];  // pre-dropout softmax
      if (keep) p[i][j] *= keep[(((size_t)b * H + h) * S + i) * S + j] * keep_scale;
    }
  }
  for (int d = lane; d < dh; d += 32) {
    float v[S];
#pragma unroll
    for (int t = 0; t < S; ++t) v[t] = qkv[((size_t)t * B + b) * 3 * E + 2 * E + h * dh + d];
#pragma unroll
    for (int i = 0; i < S; ++i) {
      float o = 0.f;
#pragma unroll
      for (int j = 0; j < S; ++j) o = fmaf(p[i][j], v[j], o);
      out[((size_t)i * B + b) * E + h * dh + d] = o;
    }
  }
}

template <int S>
__global__ void __launch_bounds__(256) token_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs,
                                                             const float* __restrict__ keep, float keep_scale,
                                                             const float* __restrict__ dout, int B, int E, int H,
                                                             float* __restrict__ dqkv) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  const int b = wid / H, h = wid - b * H;
  const int dh = E / H;
  const float scale = rsqrtf((float)dh);
  float p[S][S], pd[S][S], dpd[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i)
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const size_t o = (((size_t)b * H + h) * S + i) * S + j;
      p[i][j] = probs[o];
      const float m = keep ? keep[o] * keep_scale : 1.f;
      pd[i][j] = p[i][j] * m;  // dropped probabilities used in the forward product
      dpd[i][j] = 0.f;
    }
  // dV_j = sum_i pd_ij dO_i ;  dPd_ij = dO_i . v_j
  for (int d = lane; d < dh; d += 32) {
    float v[S], go[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
      v[t] = qkv[((size_t)t * B + b) * 3 * E + 2 * E + h * dh + d];
      go[t] = dout[((size_t)t * B + b) * E + h * dh + d];
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
      float dv = 0.f;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        dv = fmaf(pd[i][j], go[i], dv);
        dpd[i][j] = fmaf(go[i], v[j], dpd[i][j]);
      }
      dqkv[((size_t)j * B + b) * 3 * E + 2 * E + h * dh + d] = dv;
    }
  }
  float ds[S][S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float dot = 0.f;
    float dp[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
      const size_t o = (((size_t)b * H + h) * S + i) * S + j;
      const float m = keep ? keep[o] * keep_scale : 1.f;
      dp[j] = warp_sum(dpd[i][j]) * m;  // gradient wrt the softmax output
      dot = fmaf(p[i][j], dp[j], dot);
    }
#pragma unroll
    for (int j = 0; j < S; ++j) ds[i][j] = p[i][j] * (dp[j] - dot) * scale;
  }
  for (int d = lane; d < dh; d += 32) {
    float q[S], k[S];
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const float* row = qkv + ((size_t)t * B + b) * 3 * E + h * dh + d;
      q[t] = row[0];
      k[t] = row[E];
    }
#pragma unroll
    for (int t = 0; t < S; ++t) {
      float dq = 0.f, dk = 0.f;
#pragma unroll
      for (int u = 0; u < S; ++u) {
        dq = fmaf(ds[t][u], k[u], dq);
        dk = fmaf(ds[u][t], q[u], dk);
      }
      float* row = dqkv + ((size_t)t * B + b) * 3 * E + h * dh + d;
      row[0] = dq;
      row[E] = dk;
    }
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_token_attention_fwd(const float* qkv, int S, int B, int E, int H, const float* keep_mask,
                               float keep_scale, float* out, float* probs, void* stream) {
  if (!qkv || !out || !probs || S < 1 || S > kMaxTokens || B <= 0 || E <= 0 || H <= 0 || E % H != 0)
    return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (B * H * 32 + 255) / 256;
  switch (S) {
    case 1: token_attn_fwd_kernel<1><<<grid, 256, 0, s>>>(qkv, B, E, H, keep_mask, keep_scale, out, probs); break;
    case 2: token_attn_fwd_kernel<2><<<grid, 256, 0, s>>>(qkv, B, E, H, keep_mask, keep_scale, out, probs); break;
    case 3: token_attn_fwd_kernel<3><<<grid, 256, 0, s>>>(qkv, B, E, H, keep_mask, keep_scale, out, probs); break;
    default: token_attn_fwd_kernel<4><<<grid, 256, 0, s>>>(qkv, B, E, H, keep_mask, keep_scale, out, probs); break;
  }
  B200OT_LAUNCH_OK();
  return 0;
}

int b200ot_token_attention_bwd(const float* qkv, const float* probs, const float* keep_mask, float keep_scale,
                               const float* dout, int S, int B, int E, int H, float* dqkv, void* stream) {
  if (!qkv || !probs || !dout || !dqkv || S < 1 || S > kMaxTokens || B <= 0 || E <= 0 || H <= 0 || E % H != 0)
    return B200OT_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (B * H * 32 + 255) / 256;
  switch (S) {
    case 1: token_attn_bwd_kernel<1><<<grid, 256, 0, s>>>(qkv, probs, keep_mask, keep_scale, dout, B, E, H, dqkv); break;
    case 2: token_attn_bwd_kernel<2><<<grid, 256, 0, s>>>(qkv, probs, keep_mask, keep_scale, dout, B, E, H, dqkv); break;
    case 3: token_attn_bwd_kernel<3><<<grid, 256, 0, s>>>(qkv, probs, keep_mask, keep_scale, dout, B, E, H, dqkv); break;
    default: token_attn_bwd_kernel<4><<<grid, 256, 0, s>>>(qkv, probs, keep_mask, keep_scale, dout, B, E, H, dqkv); break;
  }
  B200OT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
