// AutoDis soft-bucket encoder of xDeepFM Pro's dense features.
//
// Replaces (reference, file:line), for every dense feature f of a sample with value x:
//   bucket_projectors[f] = Linear(1, nb) -> LeakyReLU(0.2) -> Linear(nb, nb)          deepctr/xdeepfm_pro/autodis.py:63-69, 107
//   softmax(scores / feature_temperatures[f])                                        deepctr/xdeepfm_pro/autodis.py:110-111
//   bucket_weights @ meta_embeddings[f]  ([nb] x [nb, E])                             deepctr/xdeepfm_pro/autodis.py:117-120
//   cat over features -> [B, nd * E]                                                  deepctr/xdeepfm_pro/autodis.py:126-127
// and torch autograd of those ops.  The reference runs 13 x (2 Linear + LeakyReLU + softmax + matmul) = ~80 launches forward and
// about twice that backward; here one launch does the forward and two the backward (per-feature partial sums over sample chunks in
// a fixed order, then a fixed-order second stage: bit-reproducible gradients).
//
// Parameters arrive packed per feature (the host side stacks the reference's per-feature tensors):
//   w1 [nd, nb], b1 [nd, nb], W2 [nd, nb, nb] (row k = output bucket), b2 [nd, nb], meta [nd, nb, E], temp [nd].
// One thread = one (sample, feature); the feature's weights sit in shared memory (broadcast reads), the thread's vectors
// (hidden, probabilities ...) live in thread-private shared-memory columns [index][thread] (conflict-free).
#include "common.cuh"
#include <math_constants.h>
#include "../../include/xdfm.h"

#define AD_THREADS 128

struct AutoDisShapes {
  int nb, E;
  int n_w;        // floats of one feature's weights in shared memory
  int n_par;      // trainable scalars per feature, gradient pack order: meta | W2 | b2 | w1 | b1 | temp
};

__host__ __device__ static inline AutoDisShapes autodis_shapes(int nb, int E) {
  AutoDisShapes s;
  s.nb = nb; s.E = E;
  s.n_w = nb * E + nb * nb + 3 * nb;
  s.n_par = s.n_w + 1;
  return s;
}

// weights of feature f -> shared memory, layout: meta [nb*E] | W2 [nb*nb] | b2 [nb] | w1 [nb] | b1 [nb]
__device__ __forceinline__ void autodis_load_weights(float* sw, int f, int nb, int E, const float* __restrict__ w1,
                                                     const float* __restrict__ b1, const float* __restrict__ W2,
                                                     const float* __restrict__ b2, const float* __restrict__ meta) {
  float* s_meta = sw;
  float* s_W2 = s_meta + nb * E;
  float* s_b2 = s_W2 + nb * nb;
  float* s_w1 = s_b2 + nb;
  float* s_b1 = s_w1 + nb;
  for (int i = threadIdx.x; i < nb * E; i += blockDim.x) s_meta[i] = meta[(int64_t)f * nb * E + i];
  for (int i = threadIdx.x; i < nb * nb; i += blockDim.x) s_W2[i] = W2[(int64_t)f * nb * nb + i];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    s_b2[i] = b2[f * nb + i];
    s_w1[i] = w1[f * nb + i];
    s_b1[i] = b1[f * nb + i];
  }
}

// hidden -> colH, softmax probabilities -> colP (both [nb] columns of this thread); returns nothing else: the raw scores are
// recomputable from p only up to a constant, so the backward keeps z = score / temp in colZ when colZ != nullptr.
__device__ __forceinline__ void autodis_probs(const float* sw, int nb, int E, float x, float temp, float* colH, float* colP, float* colZ,
                                              float* colPre) {
  const float* s_W2 = sw + nb * E;
  const float* s_b2 = s_W2 + nb * nb;
  const float* s_w1 = s_b2 + nb;
  const float* s_b1 = s_w1 + nb;
  for (int l = 0; l < nb; ++l) {
    const float pre = fmaf(s_w1[l], x, s_b1[l]);
    colH[l * AD_THREADS] = pre > 0.f ? pre : 0.2f * pre;
    if (colPre) colPre[l * AD_THREADS] = pre;
  }
  float mx = -CUDART_INF_F;
  for (int k = 0; k < nb; ++k) {
    float s = s_b2[k];
    const float* wr = s_W2 + k * nb;
    for (int l = 0; l < nb; ++l) s = fmaf(wr[l], colH[l * AD_THREADS], s);
    const float z = s / temp;
    colP[k * AD_THREADS] = z;
    if (colZ) colZ[k * AD_THREADS] = z;
    mx = fmaxf(mx, z);
  }
  float sum = 0.f;
  for (int k = 0; k < nb; ++k) {
    const float e = expf(colP[k * AD_THREADS] - mx);
    colP[k * AD_THREADS] = e;
    sum += e;
  }
  const float inv = 1.f / sum;
  for (int k = 0; k < nb; ++k) colP[k * AD_THREADS] *= inv;
}

// grid (chunks, nd); dynamic smem = (n_w + 2 * nb * AD_THREADS) floats
__global__ void __launch_bounds__(AD_THREADS) autodis_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                                 const float* __restrict__ b1, const float* __restrict__ W2,
                                                                 const float* __restrict__ b2, const float* __restrict__ meta,
                                                                 const float* __restrict__ temp, int64_t B, int nd, int nb, int E,
                                                                 float* __restrict__ out) {
  extern __shared__ float ad_smem[];
  const AutoDisShapes sh = autodis_shapes(nb, E);
  const int f = blockIdx.y;
  float* sw = ad_smem;
  float* colH = sw + sh.n_w + threadIdx.x;
  float* colP = colH + nb * AD_THREADS;
  autodis_load_weights(sw, f, nb, E, w1, b1, W2, b2, meta);
  __syncthreads();
  const float t = temp[f];
  for (int64_t b = (int64_t)blockIdx.x * AD_THREADS + threadIdx.x; b < B; b += (int64_t)gridDim.x * AD_THREADS) {
    autodis_probs(sw, nb, E, x[b * nd + f], t, colH, colP, nullptr, nullptr);
    float* o = out + (b * nd + f) * (int64_t)E;
    for (int e = 0; e < E; ++e) {
      float acc = 0.f;
      for (int k = 0; k < nb; ++k) acc = fmaf(colP[k * AD_THREADS], sw[k * E + e], acc);
      o[e] = acc;
    }
  }
}

// Backward, stage 1.  grid (chunks, nd).  For each chunk of AD_THREADS samples the threads first write their per-sample vectors
//   P [nb] probabilities, DS [nb] d loss / d score, H [nb] hidden, DPRE [nb] d loss / d pre-activation, DO [E] upstream gradient,
//   X (the value), DT (d loss / d temp), ONE (= 1)
// into shared-memory columns, then every thread sums the parameters it owns (index t, t + 128, ...) as sum_s A[s] * B[s] over the
// chunk's samples, starting at sample (t mod 128) so that the 32 lanes of a warp hit 32 different banks.  Accumulators persist
// across the block's chunks; one partial vector per block goes to `partial` [chunks, nd, n_par].  Every order is a fixed function
// of (B, nd, nb, E, grid): run-to-run bit-identical.
#define AD_MAX_OWN 40
__global__ void __launch_bounds__(AD_THREADS) autodis_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                                 const float* __restrict__ b1, const float* __restrict__ W2,
                                                                 const float* __restrict__ b2, const float* __restrict__ meta,
                                                                 const float* __restrict__ temp, const float* __restrict__ dout,
                                                                 int64_t B, int nd, int nb, int E, float* __restrict__ partial) {
  extern __shared__ float ad_smem[];
  const AutoDisShapes sh = autodis_shapes(nb, E);
  const int f = blockIdx.y;
  const int tid = threadIdx.x;
  float* sw = ad_smem;
  float* colP = sw + sh.n_w;                       // [nb][128]
  float* colZ = colP + nb * AD_THREADS;            // [nb][128]  z = score / temp
  float* colDS = colZ + nb * AD_THREADS;           // [nb][128]  dp, then d loss / d score
  float* colH = colDS + nb * AD_THREADS;           // [nb][128]
  float* colDPRE = colH + nb * AD_THREADS;         // [nb][128]  pre-activation, then its gradient
  float* colDO = colDPRE + nb * AD_THREADS;        // [E][128]
  float* colX = colDO + E * AD_THREADS;            // [128]
  float* colDT = colX + AD_THREADS;                // [128]
  float* colONE = colDT + AD_THREADS;              // [128]
  autodis_load_weights(sw, f, nb, E, w1, b1, W2, b2, meta);
  colONE[tid] = 1.f;
  const float* s_W2 = sw + nb * E;
  const float t = temp[f];
  const int n_meta = nb * E, n_W2 = nb * nb;
  float acc[AD_MAX_OWN];
#pragma unroll
  for (int j = 0; j < AD_MAX_OWN; ++j) acc[j] = 0.f;
  for (int64_t b0 = (int64_t)blockIdx.x * AD_THREADS; b0 < B; b0 += (int64_t)gridDim.x * AD_THREADS) {
    __syncthreads();                                // weights loaded / previous chunk's columns consumed
    const int64_t b = b0 + tid;
    if (b < B) {
      const float xv = x[b * nd + f];
      autodis_probs(sw, nb, E, xv, t, colH + tid, colP + tid, colZ + tid, colDPRE + tid);
      const float* g = dout + (b * nd + f) * (int64_t)E;
      for (int e = 0; e < E; ++e) colDO[e * AD_THREADS + tid] = g[e];
      float pdp = 0.f;
      for (int k = 0; k < nb; ++k) {               // dp[k] = sum_e dout[e] * meta[k, e]
        float dp = 0.f;
        for (int e = 0; e < E; ++e) dp = fmaf(colDO[e * AD_THREADS + tid], sw[k * E + e], dp);
        colDS[k * AD_THREADS + tid] = dp;
        pdp = fmaf(colP[k * AD_THREADS + tid], dp, pdp);
      }
      float dt = 0.f;
      for (int k = 0; k < nb; ++k) {               // softmax backward, then through z = score / temp
        const float dz = colP[k * AD_THREADS + tid] * (colDS[k * AD_THREADS + tid] - pdp);
        dt = fmaf(-dz, colZ[k * AD_THREADS + tid], dt);
        colDS[k * AD_THREADS + tid] = dz / t;
      }
      colDT[tid] = dt / t;
      for (int l = 0; l < nb; ++l) {               // second Linear backward to the hidden, LeakyReLU(0.2) backward
        float dh = 0.f;
        for (int k = 0; k < nb; ++k) dh = fmaf(colDS[k * AD_THREADS + tid], s_W2[k * nb + l], dh);
        const float pre = colDPRE[l * AD_THREADS + tid];
        colDPRE[l * AD_THREADS + tid] = pre > 0.f ? dh : 0.2f * dh;
      }
      colX[tid] = xv;
    } else {                                        // rows past the batch contribute zeros
      for (int k = 0; k < nb; ++k) {
        colP[k * AD_THREADS + tid] = 0.f;
        colDS[k * AD_THREADS + tid] = 0.f;
        colH[k * AD_THREADS + tid] = 0.f;
        colDPRE[k * AD_THREADS + tid] = 0.f;
      }
      for (int e = 0; e < E; ++e) colDO[e * AD_THREADS + tid] = 0.f;
      colX[tid] = 0.f;
      colDT[tid] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < AD_MAX_OWN; ++j) {
      const int q = tid + j * AD_THREADS;
      if (q < sh.n_par) {
        const float *A, *Bc;
        int r = q;
        if (r < n_meta) { A = colP + (r / E) * AD_THREADS; Bc = colDO + (r % E) * AD_THREADS; }
        else if ((r -= n_meta) < n_W2) { A = colDS + (r / nb) * AD_THREADS; Bc = colH + (r % nb) * AD_THREADS; }
        else if ((r -= n_W2) < nb) { A = colDS + r * AD_THREADS; Bc = colONE; }
        else if ((r -= nb) < nb) { A = colDPRE + r * AD_THREADS; Bc = colX; }
        else if ((r -= nb) < nb) { A = colDPRE + r * AD_THREADS; Bc = colONE; }
        else { A = colDT; Bc = colONE; }
        float a = acc[j];
#pragma unroll 4
        for (int i = 0; i < AD_THREADS; ++i) {
          const int s = (i + tid) & (AD_THREADS - 1);
          a = fmaf(A[s], Bc[s], a);
        }
        acc[j] = a;
      }
    }
  }
  float* dst = partial + ((int64_t)blockIdx.x * nd + f) * sh.n_par;
#pragma unroll
  for (int j = 0; j < AD_MAX_OWN; ++j) {
    const int q = tid + j * AD_THREADS;
    if (q < sh.n_par) dst[q] = acc[j];
  }
}

// stage 2: gpack[f, q] = sum over chunks (ascending) of partial[c, f, q]
__global__ void autodis_bwd_reduce_kernel(const float* __restrict__ partial, int chunks, int64_t n, float* __restrict__ gpack) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < chunks; ++c) a += partial[(int64_t)c * n + i];
    gpack[i] = a;
  }
}

static int autodis_check(int64_t B, int nd, int nb, int E, size_t smem) {
  XDFM_CHECK_ARG(B >= 0 && nd >= 1 && nd <= XDFM_MAX_DENSE && nb >= 1 && E >= 1, "autodis: bad shape B=%lld nd=%d nb=%d E=%d", (long long)B,
                 nd, nb, E);
  const AutoDisShapes sh = autodis_shapes(nb, E);
  if (smem > 200 * 1024 || sh.n_par > AD_MAX_OWN * AD_THREADS) {
    xdfm_set_error("autodis: num_buckets=%d x embedding_dim=%d is too large for the fused kernel (shared memory %zu B, %d parameters)", nb, E,
                   smem, sh.n_par);
    return XDFM_ERR_UNSUPPORTED;
  }
  return XDFM_OK;
}

static int autodis_chunks(int64_t B, int nd) {
  const int64_t by_batch = ceil_div64(B, AD_THREADS);
  const int64_t by_sms = std::max(1, 4 * xdfm_num_sms() / nd);
  return (int)std::max<int64_t>(1, std::min(by_batch, by_sms));
}

extern "C" int xdfm_autodis_fwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, const float* meta,
                                const float* temp, int64_t B, int nd, int nb, int E, float* out, void* stream) {
  const AutoDisShapes sh = autodis_shapes(nb, E);
  const size_t smem = ((size_t)sh.n_w + 2 * (size_t)nb * AD_THREADS) * sizeof(float);
  int rc = autodis_check(B, nd, nb, E, smem);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CUDA(cudaFuncSetAttribute(autodis_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(autodis_chunks(B, nd), nd);
  autodis_fwd_kernel<<<grid, AD_THREADS, smem, (cudaStream_t)stream>>>(x, w1, b1, W2, b2, meta, temp, B, nd, nb, E, out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int64_t xdfm_autodis_param_count(int nb, int E) { return autodis_shapes(nb, E).n_par; }

extern "C" int64_t xdfm_autodis_bwd_workspace_bytes(int64_t B, int nd, int nb, int E) {
  return (int64_t)autodis_chunks(std::max<int64_t>(B, 1), nd) * nd * autodis_shapes(nb, E).n_par * (int64_t)sizeof(float);
}

extern "C" int xdfm_autodis_bwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, const float* meta,
                                const float* temp, const float* dout, int64_t B, int nd, int nb, int E, float* gpack, void* workspace,
                                void* stream) {
  const AutoDisShapes sh = autodis_shapes(nb, E);
  const size_t smem = ((size_t)sh.n_w + (5 * (size_t)nb + E + 3) * AD_THREADS) * sizeof(float);
  int rc = autodis_check(B, nd, nb, E, smem);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)nd * sh.n_par;
  if (B == 0) {
    XDFM_CUDA(cudaMemsetAsync(gpack, 0, n * sizeof(float), st));
    return XDFM_OK;
  }
  XDFM_CHECK_ARG(workspace != nullptr, "autodis_bwd: workspace is null");
  XDFM_CUDA(cudaFuncSetAttribute(autodis_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int chunks = autodis_chunks(B, nd);
  dim3 grid(chunks, nd);
  autodis_bwd_kernel<<<grid, AD_THREADS, smem, st>>>(x, w1, b1, W2, b2, meta, temp, dout, B, nd, nb, E, (float*)workspace);
  XDFM_LAUNCH_CHECK();
  autodis_bwd_reduce_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 1024), 256, 0, st>>>((const float*)workspace, chunks, n, gpack);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
