// Supervised-Feature-Generation loss kernels of xDeepFM Pro.
//
// Replaces (reference, file:line):
//   positive mask / num_positive                                   deepctr/xdeepfm_pro/sfg_decoder.py:266-273
//   per-field F.cross_entropy(reduction='none') * mask / num_pos   deepctr/xdeepfm_pro/sfg_decoder.py:281-292
//   F.mse_loss(...).mean(-1) * mask / num_pos                      deepctr/xdeepfm_pro/sfg_decoder.py:296-305
// The reference materialises log-softmax + per-row losses per field and syncs the host (.item()) m+2 times per step; here one
// CTA per row does max / sum-exp / loss / gradient in one pass over the [V] logits and nothing is read back.
#include "common.cuh"
#include <math_constants.h>
#include "../../include/xdfm.h"

__device__ __forceinline__ float block_reduce_256(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

// row_w[b] = mask_b / num,  mask = (label == 1) and num = sum(mask) + 1e-8 (positive_only) or mask = 1, num = B
__global__ void __launch_bounds__(256) sfg_row_weights_kernel(const float* __restrict__ labels, int64_t B, int positive_only,
                                                              float* __restrict__ row_w) {
  __shared__ float red[8];
  float cnt = 0.f;
  if (positive_only)
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) cnt += (labels[b] == 1.f) ? 1.f : 0.f;
  cnt = block_reduce_256(cnt, red, false);
  const float num = positive_only ? cnt + 1e-8f : (float)B;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) row_w[b] = (positive_only ? ((labels[b] == 1.f) ? 1.f : 0.f) : 1.f) / num;
}

extern "C" int xdfm_sfg_row_weights(const float* labels, int64_t B, int positive_only, float* row_w, void* stream) {
  if (B == 0) return XDFM_OK;
  sfg_row_weights_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(labels, B, positive_only, row_w);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// row_loss[r] = w_r * (logsumexp(logits[r,:]) - logits[r, target_r]);  dlogits[r,v] = w_r * (softmax_v - [v == target_r])
__global__ void __launch_bounds__(256) masked_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                        int64_t tstride, const float* __restrict__ row_w, int V,
                                                        float* __restrict__ row_loss, float* __restrict__ dlogits) {
  __shared__ float red[8];
  const int64_t r = blockIdx.x;
  const float* lr = logits + r * V;
  const float w = row_w[r];
  float* dr = dlogits + r * V;
  if (w == 0.f) {       // masked row (label != 1): contributes nothing
    for (int v = threadIdx.x; v < V; v += blockDim.x) dr[v] = 0.f;
    if (threadIdx.x == 0) row_loss[r] = 0.f;
    return;
  }
  float mx = -CUDART_INF_F;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, lr[v]);
  mx = block_reduce_256(mx, red, true);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(lr[v] - mx);
  s = block_reduce_256(s, red, false);
  const float inv = 1.f / s;
  int t = targets[r * tstride];
  t = max(0, min(t, V - 1));
  for (int v = threadIdx.x; v < V; v += blockDim.x) dr[v] = w * (expf(lr[v] - mx) * inv - (v == t ? 1.f : 0.f));
  if (threadIdx.x == 0) row_loss[r] = w * (mx + logf(s) - lr[t]);
}

extern "C" int xdfm_masked_ce(const float* logits, const int32_t* targets, int64_t target_stride, const float* row_w, int64_t R, int V,
                              float* row_loss, float* dlogits, void* stream) {
  XDFM_CHECK_ARG(V >= 1, "masked_ce: V=%d", V);
  if (R == 0) return XDFM_OK;
  masked_ce_kernel<<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>(logits, targets, target_stride, row_w, V, row_loss, dlogits);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// row_loss[r] = w_r * mean_j (pred[r,j] - target[r,j])^2;  dpred[r,j] = w_r * 2 (pred - target) / nd
__global__ void __launch_bounds__(256) masked_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                         const float* __restrict__ row_w, int64_t R, int nd, float* __restrict__ row_loss,
                                                         float* __restrict__ dpred) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    const float w = row_w[r];
    float acc = 0.f;
    for (int j = 0; j < nd; ++j) {
      const float d = pred[r * nd + j] - target[r * nd + j];
      acc = fmaf(d, d, acc);
      dpred[r * nd + j] = w * 2.f * d / (float)nd;
    }
    row_loss[r] = w * acc / (float)nd;
  }
}

extern "C" int xdfm_masked_mse(const float* pred, const float* target, const float* row_w, int64_t R, int nd, float* row_loss, float* dpred,
                               void* stream) {
  XDFM_CHECK_ARG(nd >= 1, "masked_mse: nd=%d", nd);
  if (R == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(R, 256));
  masked_mse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, row_w, R, nd, row_loss, dpred);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
