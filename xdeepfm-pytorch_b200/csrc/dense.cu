// Dense helper kernels of the fp32 path: tiled SGEMM with fused bias/activation epilogue (DNN layers, attention
// projections), the logit head (linear + CIN + DNN terms, bias, sigmoid), BCE-on-probabilities forward/backward,
// deterministic (weighted) column sums for bias / GEMV-shaped weight gradients.
//
// Replaces (reference, file:line):
//   DNN.forward  (nn.Linear -> ReLU)                        deepctr/layers/core.py:120-134
//   dnn_linear / cin_linear (Linear(.,1,bias=False))         deepctr/models/xdeepfm.py:56, 73, 88, 92
//   final_logit sum + PredictionLayer (bias, sigmoid)        deepctr/models/xdeepfm.py:94-105, deepctr/layers/core.py:154-160
//   F.binary_cross_entropy(y_pred, y, reduction='sum')       deepctr/models/basemodel.py:254
#include "common.cuh"
#include "../../include/xdfm.h"

// ------------------------------------------------------------------------------------------------
// C[M,N] = act( op(A)[M,K] * op(B)[K,N] + bias[N] ),  optional C += (accumulate)
// op(A)(i,k) = transA ? A[k*lda + i] : A[i*lda + k];  op(B)(k,j) = transB ? B[j*ldb + k] : B[k*ldb + j]
// 64x64x16 tiles, 256 threads, 4x4 micro-tile.  Split-K over gridDim.z writes partials (deterministic 2nd stage).
// ------------------------------------------------------------------------------------------------
#define GT_M 64
#define GT_N 64
#define GT_K 16

__global__ void __launch_bounds__(256) sgemm_kernel(int transA, int transB, int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                    const float* __restrict__ bias, int act, int accumulate, int k_per_split,
                                                    float* __restrict__ partial) {
  __shared__ float As[GT_K][GT_M + 4];
  __shared__ float Bs[GT_K][GT_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // the larger tile count rides on gridDim.x (limit 2^31-1): M = B*L reaches 4 M rows in the attention block
  const int m_on_x = (M + GT_M - 1) / GT_M >= (N + GT_N - 1) / GT_N;
  const int m0 = (m_on_x ? blockIdx.x : blockIdx.y) * GT_M, n0 = (m_on_x ? blockIdx.y : blockIdx.x) * GT_N;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GT_K) {
    // load A tile: 64x16 = 1024 elements, 4 per thread
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;
      int i, k;
      if (transA) { i = e & 63; k = e >> 6; } else { k = e & 15; i = e >> 4; }
      int gi = m0 + i, gk = k0 + k;
      float v = 0.f;
      if (gi < M && gk < kend) v = transA ? A[(int64_t)gk * lda + gi] : A[(int64_t)gi * lda + gk];
      As[k][i] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;
      int j, k;
      if (transB) { k = e & 15; j = e >> 4; } else { j = e & 63; k = e >> 6; }
      int gj = n0 + j, gk = k0 + k;
      float v = 0.f;
      if (gj < N && gk < kend) v = transB ? B[(int64_t)gj * ldb + gk] : B[(int64_t)gk * ldb + gj];
      Bs[k][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GT_K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gi = m0 + ty * 4 + i;
    if (gi >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gj = n0 + tx * 4 + j;
      if (gj >= N) continue;
      if (partial != nullptr) {
        partial[((int64_t)blockIdx.z * M + gi) * N + gj] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (bias) v += bias[gj];
        if (act == XDFM_ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == XDFM_ACT_TANH) v = tanhf(v);
        else if (act == XDFM_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
        if (accumulate) v += C[(int64_t)gi * ldc + gj];
        C[(int64_t)gi * ldc + gj] = v;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int S, int M, int N, float* __restrict__ C, int ldc,
                                     const float* __restrict__ bias, int act, int accumulate) {
  int64_t total = (int64_t)M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int gi = (int)(i / N), gj = (int)(i - (int64_t)gi * N);
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += partial[(int64_t)s * total + i];
    if (bias) v += bias[gj];
    if (act == XDFM_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == XDFM_ACT_TANH) v = tanhf(v);
    else if (act == XDFM_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    if (accumulate) v += C[(int64_t)gi * ldc + gj];
    C[(int64_t)gi * ldc + gj] = v;
  }
}

extern "C" int64_t xdfm_gemm_workspace_bytes(int M, int N, int K) {
  // split-K only pays for small outputs with a long reduction (weight gradients)
  int64_t tiles = ceil_div64(M, GT_M) * ceil_div64(N, GT_N);
  int S = 1;
  if (tiles < 2 * xdfm_num_sms() && K >= 1024) S = (int)std::min<int64_t>(32, std::max<int64_t>(1, (2 * xdfm_num_sms()) / tiles));
  while (S > 1 && ceil_div64(K, S) < 256) --S;
  return S > 1 ? (int64_t)S * M * N * 4 : 0;
}

extern "C" int xdfm_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                             int ldc, const float* bias, int act, int accumulate, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (M == 0 || N == 0) return XDFM_OK;
  XDFM_CHECK_ARG(M > 0 && N > 0 && K >= 0, "gemm_f32: bad shape %d %d %d", M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t need = xdfm_gemm_workspace_bytes(M, N, K);
  int S = 1;
  if (need > 0 && workspace != nullptr && workspace_bytes >= need) S = (int)(need / ((int64_t)M * N * 4));
  int kps = (int)ceil_div64(max(K, 1), S);
  kps = (int)ceil_div64(kps, GT_K) * GT_K;
  S = (int)ceil_div64(max(K, 1), kps);
  const unsigned mt = (unsigned)ceil_div64(M, GT_M), nt = (unsigned)ceil_div64(N, GT_N);
  dim3 grid(mt >= nt ? mt : nt, mt >= nt ? nt : mt, (unsigned)S);
  XDFM_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_f32: shape %d x %d x %d exceeds the launch grid", M, N, K);
  if (S == 1) {
    sgemm_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, kps, nullptr);
    XDFM_LAUNCH_CHECK();
  } else {
    sgemm_kernel<<<grid, 256, 0, st>>>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, act, accumulate, kps, (float*)workspace);
    XDFM_LAUNCH_CHECK();
    int blocks = (int)min((int64_t)xdfm_num_sms() * 4, ceil_div64((int64_t)M * N, 256));
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>((const float*)workspace, S, M, N, C, ldc, bias, act, accumulate);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// elementwise: dx = dy * act'(y) given the activation OUTPUT y
// ------------------------------------------------------------------------------------------------
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = dy[i], v = y[i];
    if (act == XDFM_ACT_RELU) g = v > 0.f ? g : 0.f;
    else if (act == XDFM_ACT_TANH) g = g * (1.f - v * v);
    else if (act == XDFM_ACT_SIGMOID) g = g * v * (1.f - v);
    dx[i] = g;
  }
}

extern "C" int xdfm_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, void* stream) {
  if (n == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  act_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, act);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// deterministic weighted column sum: out[k] = sum_b s[b] * X[b*ld + k]   (s == null -> plain column sum)
// stage 1: block c handles rows [c*rows_per, ...), thread per column; stage 2: fixed-order sum over chunks.
// ------------------------------------------------------------------------------------------------
#define WCS_CHUNKS 64
__global__ void wcolsum_stage1(const float* __restrict__ X, int64_t B, int K, int ld, const float* __restrict__ s, float* __restrict__ part) {
  int c = blockIdx.y;
  int64_t rows_per = (B + WCS_CHUNKS - 1) / WCS_CHUNKS;
  int64_t b0 = c * rows_per, b1 = min(B, b0 + rows_per);
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  // eight rows in flight per thread (the plain loop paid one memory latency per row); four accumulators combined in a fixed order
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t b = b0;
  for (; b + 8 <= b1; b += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (s ? s[b + u] : 1.f) * X[(b + u) * ld + k];
    a0 += v[0]; a1 += v[1]; a2 += v[2]; a3 += v[3];
    a0 += v[4]; a1 += v[5]; a2 += v[6]; a3 += v[7];
  }
  for (; b < b1; ++b) a0 += (s ? s[b] : 1.f) * X[b * ld + k];
  part[(int64_t)c * K + k] = (a0 + a1) + (a2 + a3);
}
extern "C" int64_t xdfm_wcolsum_workspace_bytes(int K) { return (int64_t)WCS_CHUNKS * K * 4; }

extern "C" int xdfm_wcolsum(const float* X, int64_t B, int K, int ld, const float* s, float* out, int accumulate, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  if (K == 0) return XDFM_OK;
  XDFM_CHECK_ARG(workspace_bytes >= xdfm_wcolsum_workspace_bytes(K), "wcolsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g1((unsigned)ceil_div64(K, 128), WCS_CHUNKS);
  wcolsum_stage1<<<g1, 128, 0, st>>>(X, B, K, ld, s, (float*)workspace);
  XDFM_LAUNCH_CHECK();
  XDFM_TILE_COLSUM((const float*)workspace, (int64_t)WCS_CHUNKS, (int64_t)K, K, out, accumulate, st);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// logit head: y = sigmoid(lin + <cin_out, w_cin> + <dnn_out, w_dnn> + bias); one warp per sample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ lin, const float* __restrict__ cin_out,
                                                       const float* __restrict__ w_cin, int fm, const float* __restrict__ dnn_out,
                                                       const float* __restrict__ w_dnn, int hd, const float* __restrict__ bias,
                                                       int64_t B, int binary, float* __restrict__ y_pred) {
  int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float c = 0.f, d = 0.f;
    if (cin_out) for (int k = lane; k < fm; k += 32) c = fmaf(cin_out[b * fm + k], w_cin[k], c);
    if (dnn_out) for (int k = lane; k < hd; k += 32) d = fmaf(dnn_out[b * hd + k], w_dnn[k], d);
    c = warp_sum(c);
    d = warp_sum(d);
    if (lane == 0) {
      // reference order: linear_logit + dnn_logit + cin_logit, then += bias (xdeepfm.py:101, core.py:157)
      float z = (lin ? lin[b] : 0.f);
      if (dnn_out) z += d;
      if (cin_out) z += c;
      if (bias) z += bias[0];
      y_pred[b] = binary ? 1.f / (1.f + expf(-z)) : z;
    }
  }
}

extern "C" int xdfm_head_fwd(const float* lin, const float* cin_out, const float* w_cin, int fm, const float* dnn_out,
                             const float* w_dnn, int hd, const float* bias, int64_t B, int binary, float* y_pred, void* stream) {
  if (B == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(B, 8));
  head_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lin, cin_out, w_cin, fm, dnn_out, w_dnn, hd, bias, B, binary, y_pred);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// dlogit = dy_pred * p(1-p) (binary) or dy_pred; then d_cin_out[b,:] = dlogit[b]*w_cin, d_dnn_out[b,:] = dlogit[b]*w_dnn
__global__ void head_bwd_kernel(const float* __restrict__ dy_pred, const float* __restrict__ y_pred, int64_t B, int binary,
                                const float* __restrict__ w_cin, int fm, const float* __restrict__ w_dnn, int hd,
                                float* __restrict__ dlogit, float* __restrict__ d_cin_out, float* __restrict__ d_dnn_out) {
  int W = fm + hd + 1;
  int64_t total = B * (int64_t)W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / W;
    int c = (int)(i - b * W);
    float g = dy_pred[b];
    if (binary) { float p = y_pred[b]; g = g * (p * (1.f - p)); }
    if (c == 0) dlogit[b] = g;
    else if (c <= fm) { if (d_cin_out) d_cin_out[b * fm + (c - 1)] = g * w_cin[c - 1]; }
    else { if (d_dnn_out) d_dnn_out[b * hd + (c - 1 - fm)] = g * w_dnn[c - 1 - fm]; }
  }
}

extern "C" int xdfm_head_bwd(const float* dy_pred, const float* y_pred, int64_t B, int binary, const float* w_cin, int fm,
                             const float* w_dnn, int hd, float* dlogit, float* d_cin_out, float* d_dnn_out, void* stream) {
  if (B == 0) return XDFM_OK;
  if (!d_cin_out) fm = 0;
  if (!d_dnn_out) hd = 0;
  int64_t total = B * (int64_t)(fm + hd + 1);
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  head_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy_pred, y_pred, B, binary, w_cin, fm, w_dnn, hd, dlogit, d_cin_out,
                                                            d_dnn_out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// Fused backward of the head: one pass over the saved CIN / DNN outputs gives dlogit, d_cin_out, d_dnn_out AND the per-chunk partial
// sums of d_w_cin = sum_b dlogit[b] * cin_out[b, :], d_w_dnn, d_bias = sum_b dlogit[b] (xdfm_head_bwd + three xdfm_wcolsum = 7 launches
// before).  Block = 16 samples; thread = a column of [cin_out | dnn_out]; part [n_chunks, fm + hd + 1]; fixed order throughout.
#define HEAD_ROWS 16
__global__ void __launch_bounds__(256) head_bwd_fused_kernel(const float* __restrict__ dy_pred, const float* __restrict__ y_pred, int64_t B,
                                                             int binary, const float* __restrict__ cin_out, const float* __restrict__ w_cin,
                                                             int fm, const float* __restrict__ dnn_out, const float* __restrict__ w_dnn, int hd,
                                                             float* __restrict__ dlogit, float* __restrict__ d_cin_out,
                                                             float* __restrict__ d_dnn_out, float* __restrict__ part) {
  __shared__ float sg[HEAD_ROWS];
  const int64_t b0 = (int64_t)blockIdx.x * HEAD_ROWS;
  const int W = fm + hd;
  if (threadIdx.x < HEAD_ROWS) {
    const int64_t b = b0 + threadIdx.x;
    float g = 0.f;
    if (b < B) {
      g = dy_pred[b];
      if (binary) { const float p = y_pred[b]; g = g * (p * (1.f - p)); }
      dlogit[b] = g;
    }
    sg[threadIdx.x] = g;                       // rows past B contribute zeros
  }
  __syncthreads();
  float* prow = part + (int64_t)blockIdx.x * (W + 1);
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_ROWS; ++i) a += sg[i];
    prow[W] = a;
  }
  const int nrow = (int)min((int64_t)HEAD_ROWS, B - b0);
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    const bool in_cin = c < fm;
    const int cc = in_cin ? c : c - fm;
    const int ld = in_cin ? fm : hd;
    const float* src = (in_cin ? cin_out : dnn_out) + b0 * ld + cc;
    float* dst = (in_cin ? d_cin_out : d_dnn_out) + b0 * ld + cc;
    const float w = in_cin ? w_cin[cc] : w_dnn[cc];
    float v[HEAD_ROWS];
#pragma unroll
    for (int i = 0; i < HEAD_ROWS; ++i) v[i] = i < nrow ? src[(int64_t)i * ld] : 0.f;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_ROWS; i += 2) {
      a0 = fmaf(sg[i], v[i], a0);
      a1 = fmaf(sg[i + 1], v[i + 1], a1);
    }
#pragma unroll
    for (int i = 0; i < HEAD_ROWS; ++i)
      if (i < nrow) dst[(int64_t)i * ld] = sg[i] * w;
    prow[c] = a0 + a1;
  }
}

extern "C" int64_t xdfm_head_bwd_fused_workspace_bytes(int64_t B, int fm, int hd) {
  return ceil_div64(B, HEAD_ROWS) * (int64_t)(fm + hd + 1) * 4;
}

// d_w [fm + hd + 1] = (d_w_cin | d_w_dnn | d_bias).  cin_out / dnn_out may be NULL (then fm / hd count as 0).
extern "C" int xdfm_head_bwd_fused(const float* dy_pred, const float* y_pred, int64_t B, int binary, const float* cin_out, const float* w_cin,
                                   int fm, const float* dnn_out, const float* w_dnn, int hd, float* dlogit, float* d_cin_out,
                                   float* d_dnn_out, float* d_w, void* workspace, int64_t workspace_bytes, void* stream) {
  if (cin_out == nullptr) fm = 0;
  if (dnn_out == nullptr) hd = 0;
  XDFM_CHECK_ARG(fm == 0 || (w_cin != nullptr && d_cin_out != nullptr), "head_bwd_fused: w_cin / d_cin_out missing");
  XDFM_CHECK_ARG(hd == 0 || (w_dnn != nullptr && d_dnn_out != nullptr), "head_bwd_fused: w_dnn / d_dnn_out missing");
  XDFM_CHECK_ARG(d_w != nullptr && dlogit != nullptr, "head_bwd_fused: d_w / dlogit missing");
  XDFM_CHECK_ARG(workspace != nullptr && workspace_bytes >= xdfm_head_bwd_fused_workspace_bytes(B, fm, hd), "head_bwd_fused: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int W = fm + hd;
  if (B == 0) {
    XDFM_CUDA(cudaMemsetAsync(d_w, 0, (size_t)(W + 1) * 4, st));
    return XDFM_OK;
  }
  const int64_t n_chunks = ceil_div64(B, HEAD_ROWS);
  head_bwd_fused_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(dy_pred, y_pred, B, binary, cin_out, w_cin, fm, dnn_out, w_dnn, hd, dlogit,
                                                          d_cin_out, d_dnn_out, (float*)workspace);
  XDFM_LAUNCH_CHECK();
  XDFM_TILE_COLSUM((const float*)workspace, n_chunks, (int64_t)(W + 1), W + 1, d_w, 0, st);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// BCE on probabilities, reduction='sum' (torch semantics: log terms clamped at -100; grad denominator clamped 1e-12)
//   loss_sum += sum_b -(y*max(log p,-100) + (1-y)*max(log(1-p),-100));  dy_pred[b] = scale*(p-y)/max(p(1-p),1e-12)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t B, float scale,
                                                  float* __restrict__ per_sample, float* __restrict__ dy_pred, double* loss_sum) {
  float local = 0.f;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float pb = p[b], yb = y[b];
    float l1 = fmaxf(logf(pb), -100.f), l0 = fmaxf(log1pf(-pb), -100.f);
    float l = -(yb * l1 + (1.f - yb) * l0);
    if (per_sample) per_sample[b] = l;
    local += l;
    if (dy_pred) dy_pred[b] = scale * (pb - yb) / fmaxf(pb * (1.f - pb), 1e-12f);
  }
  __shared__ float red[8];
  float v = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0 && loss_sum != nullptr) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)red[i];
    atomicAdd(loss_sum, t);
  }
}

extern "C" int xdfm_bce_sum(const float* y_pred, const float* labels, int64_t B, float scale, float* per_sample, float* dy_pred,
                            double* loss_sum, void* stream) {
  if (B == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 4, ceil_div64(B, 256));
  bce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y_pred, labels, B, scale, per_sample, dy_pred, loss_sum);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
