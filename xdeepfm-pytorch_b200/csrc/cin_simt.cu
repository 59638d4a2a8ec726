// CIN layer, fp32 CUDA-core path (exact-precision mode; the bf16 tcgen05 path lives in cin_tc.cu).
//
// Replaces (reference, file:line): deepctr/layers/interaction.py:207-248 -- per layer
//   Z = einsum('bhd,bmd->bhmd', X^{k-1}, X^0).reshape(B, h*m, D)   (never materialised here: generated in smem)
//   Y = relu(Conv1d_k=1(Z)) = relu(W[H,K] Z[K,(b,d)] + bias)        (implicit GEMM, K index = i*m + j)
//   split-half -> next_hidden = Y[:, :H/2], direct = Y[:, H/2:];  result = cat(directs).sum(-1)
// and its autograd backward (dW, db, dX^{k-1}, dX^0).
//
// Layouts: x0 [B, m, D]; xk = layer input [B, Hp, D] given by pointer + batch stride (for k>0 it is the first Hp
// channels of the previous layer's y); W [H, Hp*m] (the reference's Conv1d weight [H, K, 1] as is); y [B, H, D].
#include "common.cuh"
#include "../../include/xdfm.h"

#define CK 16  // K-chunk of the implicit GEMM

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int N_T>
__global__ void __launch_bounds__(256) cin_fwd_f32_kernel(const float* __restrict__ x0, const float* __restrict__ xk, int64_t xk_bstride,
                                                          const float* __restrict__ W, const float* __restrict__ bias, int64_t B, int m,
                                                          int Hp, int H, int D, int act, float* __restrict__ y, int hdb,
                                                          float* __restrict__ pooled, float* __restrict__ maps, int fm_total, int col_off) {
  constexpr int TXN = N_T / 4;         // threads along n
  constexpr int TYH = 256 / TXN;       // threads along h
  constexpr int H_BLK = TYH * 8;       // h rows per pass
  extern __shared__ float smem[];
  const int TB = N_T / D;              // samples per tile
  const int n_used = TB * D;
  float* x0s = smem;                   // [m][N_T]
  float* xks = x0s + m * N_T;          // [Hp][N_T]
  float* Ws = xks + Hp * N_T;          // [CK][H_BLK]
  float* Zs = Ws + CK * H_BLK;         // [CK][N_T]
  float* Ys = Zs + CK * N_T;           // [H_BLK][N_T + 1]
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int K = Hp * m;
  const int64_t n_tiles = (B + TB - 1) / TB;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * TB;
    __syncthreads();
    for (int e = tid; e < (m + Hp) * N_T; e += 256) {
      int r = e / N_T, n = e - r * N_T;
      int bl = n / D, d = n - bl * D;
      float v = 0.f;
      if (n < n_used && b0 + bl < B) {
        v = (r < m) ? x0[((b0 + bl) * m + r) * (int64_t)D + d] : xk[(b0 + bl) * xk_bstride + (int64_t)(r - m) * D + d];
      }
      smem[e] = v;
    }
    __syncthreads();
    for (int h0 = 0; h0 < H; h0 += H_BLK) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int k0 = 0; k0 < K; k0 += CK) {
        for (int e = tid; e < CK * H_BLK; e += 256) {
          int kk = e % CK, hh = e / CK;
          int h = h0 + hh, k = k0 + kk;
          Ws[kk * H_BLK + hh] = (h < H && k < K) ? W[(int64_t)h * K + k] : 0.f;
        }
        for (int e = tid; e < CK * N_T; e += 256) {
          int kk = e / N_T, n = e - kk * N_T;
          int k = k0 + kk;
          float z = 0.f;
          if (k < K) {
            int i = k / m, j = k - i * m;
            z = xks[i * N_T + n] * x0s[j * N_T + n];
          }
          Zs[e] = z;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CK; ++kk) {
          float a[8], b[4];
          *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(Ws + kk * H_BLK + ty * 8);
          *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(Ws + kk * H_BLK + ty * 8 + 4);
          *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(Zs + kk * N_T + tx * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
      // epilogue: bias + activation -> smem tile -> global y / maps / pooled
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int hh = ty * 8 + i, h = h0 + hh;
        float bv = (h < H) ? bias[h] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][j] + bv;
          if (act == XDFM_ACT_RELU) v = fmaxf(v, 0.f);
          else if (act == XDFM_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
          Ys[hh * (N_T + 1) + tx * 4 + j] = v;
        }
      }
      __syncthreads();
      for (int e = tid; e < H_BLK * N_T; e += 256) {
        int hh = e / N_T, n = e - hh * N_T;
        int h = h0 + hh;
        int bl = n / D, d = n - bl * D;
        if (h < H && n < n_used && b0 + bl < B) {
          float v = Ys[hh * (N_T + 1) + n];
          y[((b0 + bl) * H + h) * (int64_t)D + d] = v;
          if (maps != nullptr && h >= hdb) maps[((b0 + bl) * fm_total + col_off + (h - hdb)) * (int64_t)D + d] = v;
        }
      }
      if (pooled != nullptr) {
        for (int e = tid; e < H_BLK * TB; e += 256) {
          int hh = e / TB, bl = e - hh * TB;
          int h = h0 + hh;
          if (h < H && h >= hdb && b0 + bl < B) {
            float s = 0.f;
            for (int d = 0; d < D; ++d) s += Ys[hh * (N_T + 1) + bl * D + d];
            pooled[(b0 + bl) * fm_total + col_off + (h - hdb)] = s;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <int N_T>
static size_t cin_fwd_smem(int m, int Hp) {
  constexpr int H_BLK = (256 / (N_T / 4)) * 8;
  return sizeof(float) * ((size_t)(m + Hp) * N_T + CK * H_BLK + CK * N_T + (size_t)H_BLK * (N_T + 1));
}

extern "C" int xdfm_cin_fwd_f32(const float* x0, const float* xk, int64_t xk_bstride, const float* W, const float* bias, int64_t B, int m,
                                int Hp, int H, int D, int act, float* y, int direct_begin, float* pooled, float* maps, int fm_total,
                                int col_off, void* stream) {
  XDFM_CHECK_ARG(D >= 1 && D <= 128, "cin_fwd_f32: D=%d unsupported (1..128)", D);
  XDFM_CHECK_ARG(m >= 1 && Hp >= 1 && H >= 1, "cin_fwd_f32: bad sizes m=%d Hp=%d H=%d", m, Hp, H);
  if (B == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t limit = 200 * 1024;
#define CIN_FWD_LAUNCH(NT)                                                                                                   \
  {                                                                                                                          \
    size_t sm = cin_fwd_smem<NT>(m, Hp);                                                                                     \
    XDFM_CUDA(cudaFuncSetAttribute(cin_fwd_f32_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));          \
    int TB = NT / D;                                                                                                         \
    int64_t tiles = ceil_div64(B, TB);                                                                                       \
    int blocks = (int)std::min<int64_t>(tiles, (int64_t)xdfm_num_sms() * 2);                                                      \
    cin_fwd_f32_kernel<NT><<<blocks, 256, sm, st>>>(x0, xk, xk_bstride, W, bias, B, m, Hp, H, D, act, y, direct_begin, pooled, \
                                                    maps, fm_total, col_off);                                                \
  }
  if (D <= 128 && cin_fwd_smem<128>(m, Hp) <= limit) CIN_FWD_LAUNCH(128)
  else if (D <= 64 && cin_fwd_smem<64>(m, Hp) <= limit) CIN_FWD_LAUNCH(64)
  else if (D <= 32 && cin_fwd_smem<32>(m, Hp) <= limit) CIN_FWD_LAUNCH(32)
  else {
    xdfm_set_error("cin_fwd_f32: m=%d Hp=%d D=%d does not fit shared memory", m, Hp, D);
    return XDFM_ERR_UNSUPPORTED;
  }
#undef CIN_FWD_LAUNCH
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// backward step 1: dy = act'(y) * (direct-path grad + next-layer grad)
//   direct channels [hdb, H): dpooled[b, col_off + h - hdb] (broadcast over d) or dmaps[b, col_off + h - hdb, d]
//   next channels   [0, n_next): dnext[b, h, d]  (= dxk of layer k+1)
// ------------------------------------------------------------------------------------------------
__global__ void cin_dy_kernel(const float* __restrict__ y, int64_t B, int H, int D, int act, int hdb, const float* __restrict__ dpooled,
                              const float* __restrict__ dmaps, int fm_total, int col_off, const float* __restrict__ dnext, int n_next,
                              float* __restrict__ dy) {
  int64_t total = B * (int64_t)H * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(e % D);
    int64_t bh = e / D;
    int h = (int)(bh % H);
    int64_t b = bh / H;
    float g = 0.f;
    if (h >= hdb) {
      if (dmaps) g += dmaps[(b * fm_total + col_off + (h - hdb)) * (int64_t)D + d];
      if (dpooled) g += dpooled[b * fm_total + col_off + (h - hdb)];
    }
    if (h < n_next && dnext) g += dnext[(b * n_next + h) * (int64_t)D + d];
    float v = y[e];
    if (act == XDFM_ACT_RELU) g = v > 0.f ? g : 0.f;
    else if (act == XDFM_ACT_SIGMOID) g = g * v * (1.f - v);
    dy[e] = g;
  }
}

extern "C" int xdfm_cin_dy(const float* y, int64_t B, int H, int D, int act, int direct_begin, const float* dpooled, const float* dmaps,
                           int fm_total, int col_off, const float* dnext, int n_next, float* dy, void* stream) {
  int64_t total = B * (int64_t)H * D;
  if (total == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  cin_dy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, B, H, D, act, direct_begin, dpooled, dmaps, fm_total, col_off, dnext, n_next, dy);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// db[h] = sum_{b,d} dy[b,h,d]: deterministic two-stage (chunks of samples, fixed order)
#define DB_CHUNKS 64
__global__ void cin_db_stage1(const float* __restrict__ dy, int64_t B, int H, int D, float* __restrict__ part) {
  int c = blockIdx.y;
  int64_t per = (B + DB_CHUNKS - 1) / DB_CHUNKS;
  int64_t b0 = c * per, b1 = min(B, b0 + per);
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float acc = 0.f;
  for (int64_t b = b0; b < b1; ++b)
    for (int d = 0; d < D; ++d) acc += dy[(b * H + h) * (int64_t)D + d];
  part[c * H + h] = acc;
}
__global__ void cin_db_stage2(const float* __restrict__ part, int H, float* __restrict__ db) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  float acc = 0.f;
  for (int c = 0; c < DB_CHUNKS; ++c) acc += part[c * H + h];
  db[h] = acc;
}

// ------------------------------------------------------------------------------------------------
// backward step 2: dW[h,k] = sum_{b,d} dy[b,h,d] * xk[b,i,d] * x0[b,j,d]   (k = i*m + j)
// 64(h) x 64(k) output tiles, reduction over n = b*D + d in chunks of 16, split over gridDim.z sample ranges
// into partial buffers that are summed in fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cin_dw_f32_kernel(const float* __restrict__ x0, const float* __restrict__ xk, int64_t xk_bstride,
                                                         const float* __restrict__ dy, int64_t B, int m, int Hp, int H, int D,
                                                         int64_t samples_per_split, float* __restrict__ part) {
  __shared__ float As[CK][64 + 4];   // dy chunk  [n][h]
  __shared__ float Zs[CK][64 + 4];   // z chunk   [n][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int K = Hp * m;
  const int h0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int64_t bbeg = blockIdx.z * samples_per_split;
  const int64_t bend = min(B, bbeg + samples_per_split);
  const int64_t nbeg = bbeg * D, nend = bend * D;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t n0 = nbeg; n0 < nend; n0 += CK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256;
      int nn = e & 15, c = e >> 4;           // nn fastest -> contiguous d
      int64_t n = n0 + nn;
      float a = 0.f, z = 0.f;
      if (n < nend) {
        int64_t b = n / D;
        int d = (int)(n - b * D);
        int h = h0 + c, k = k0 + c;
        if (h < H) a = dy[(b * H + h) * (int64_t)D + d];
        if (k < K) {
          int i = k / m, j = k - i * m;
          z = xk[b * xk_bstride + (int64_t)i * D + d] * x0[(b * m + j) * (int64_t)D + d];
        }
      }
      As[nn][c] = a;
      Zs[nn][c] = z;
    }
    __syncthreads();
#pragma unroll
    for (int nn = 0; nn < CK; ++nn) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[nn][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Zs[nn][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int h = h0 + ty * 4 + i;
    if (h >= H) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < K) part[((int64_t)blockIdx.z * H + h) * K + k] = acc[i][j];
    }
  }
}

__global__ void cin_dw_reduce_kernel(const float* __restrict__ part, int S, int64_t HK, float* __restrict__ dW) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < HK; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part[(int64_t)s * HK + i];
    dW[i] = v;
  }
}

static int cin_dw_splits(int64_t B, int H, int K) {
  int64_t tiles = ceil_div64(H, 64) * ceil_div64(K, 64);
  int64_t S = std::max<int64_t>(1, (4 * (int64_t)xdfm_num_sms()) / tiles);
  S = std::min<int64_t>(S, 32);
  S = std::min<int64_t>(S, std::max<int64_t>(1, B / 16));
  return (int)S;
}

extern "C" int64_t xdfm_cin_bwd_f32_workspace_bytes(int64_t B, int m, int Hp, int H, int D) {
  int K = Hp * m;
  int S = cin_dw_splits(B, H, K);
  return (int64_t)S * H * K * 4 + (int64_t)DB_CHUNKS * H * 4 + 256;
}

// ------------------------------------------------------------------------------------------------
// backward step 3: dz[k,n] = sum_h W[h,k] dy[h,n];  dxk[i,n] = sum_j dz[(i,j),n] x0[j,n];  dx0[j,n] += sum_i dz[(i,j),n] xk[i,n]
// One block owns a tile of samples (N_T columns) -> all updates of dxk / dx0 for those samples are race-free and in fixed order.
// ------------------------------------------------------------------------------------------------
#define DX_KC 64
template <int N_T>
__global__ void __launch_bounds__(256) cin_dx_f32_kernel(const float* __restrict__ x0, const float* __restrict__ xk, int64_t xk_bstride,
                                                         const float* __restrict__ W, const float* __restrict__ dy, int64_t B, int m, int Hp,
                                                         int H, int D, float* __restrict__ dxk, float* __restrict__ dx0) {
  constexpr int TXN = N_T / 4;
  constexpr int TYK = 256 / TXN;         // threads along k
  constexpr int KPT = DX_KC / TYK;       // k rows per thread
  extern __shared__ float smem[];
  const int TB = N_T / D;
  const int n_used = TB * D;
  float* x0s = smem;                     // [m][N_T]
  float* dx0s = x0s + m * N_T;           // [m][N_T]
  float* Wc = dx0s + m * N_T;            // [CK][DX_KC]   W chunk  [h][k]
  float* dyc = Wc + CK * DX_KC;          // [CK][N_T]     dy chunk [h][n]
  float* dzs = dyc + CK * N_T;           // [DX_KC][N_T]
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int K = Hp * m;
  const int64_t n_tiles = (B + TB - 1) / TB;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b0 = tile * TB;
    __syncthreads();
    for (int e = tid; e < m * N_T; e += 256) {
      int r = e / N_T, n = e - r * N_T;
      int bl = n / D, d = n - bl * D;
      x0s[e] = (n < n_used && b0 + bl < B) ? x0[((b0 + bl) * m + r) * (int64_t)D + d] : 0.f;
      dx0s[e] = 0.f;
    }
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += DX_KC) {
      float acc[KPT][4];
#pragma unroll
      for (int i = 0; i < KPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int h0 = 0; h0 < H; h0 += CK) {
        for (int e = tid; e < CK * DX_KC; e += 256) {
          int kk = e % DX_KC, hh = e / DX_KC;
          int h = h0 + hh, k = k0 + kk;
          Wc[e] = (h < H && k < K) ? W[(int64_t)h * K + k] : 0.f;
        }
        for (int e = tid; e < CK * N_T; e += 256) {
          int hh = e / N_T, n = e - hh * N_T;
          int h = h0 + hh;
          int bl = n / D, d = n - bl * D;
          dyc[e] = (h < H && n < n_used && b0 + bl < B) ? dy[((b0 + bl) * H + h) * (int64_t)D + d] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int hh = 0; hh < CK; ++hh) {
          float a[KPT], b[4];
#pragma unroll
          for (int i = 0; i < KPT; ++i) a[i] = Wc[hh * DX_KC + ty * KPT + i];
          *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(dyc + hh * N_T + tx * 4);
#pragma unroll
          for (int i = 0; i < KPT; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < KPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dzs[(ty * KPT + i) * N_T + tx * 4 + j] = acc[i][j];
      __syncthreads();
      // contraction of the dz chunk: column n is owned by thread n (dxk) and thread N_T + n (dx0)
      const int kend = min(DX_KC, K - k0);
      if (tid < N_T) {
        int n = tid;
        int bl = n / D, d = n - bl * D;
        if (n < n_used && b0 + bl < B) {
          int i_cur = k0 / m;
          float a = 0.f;
          for (int kk = 0; kk < kend; ++kk) {
            int k = k0 + kk;
            int i = k / m, j = k - i * m;
            if (i != i_cur) {
              dxk[((b0 + bl) * Hp + i_cur) * (int64_t)D + d] += a;
              a = 0.f;
              i_cur = i;
            }
            a = fmaf(dzs[kk * N_T + n], x0s[j * N_T + n], a);
          }
          dxk[((b0 + bl) * Hp + i_cur) * (int64_t)D + d] += a;
        }
      } else if (tid < 2 * N_T) {
        int n = tid - N_T;
        int bl = n / D, d = n - bl * D;
        if (n < n_used && b0 + bl < B) {
          const float* xkp = xk + (b0 + bl) * xk_bstride + d;
          for (int kk = 0; kk < kend; ++kk) {
            int k = k0 + kk;
            int i = k / m, j = k - i * m;
            dx0s[j * N_T + n] = fmaf(dzs[kk * N_T + n], xkp[(int64_t)i * D], dx0s[j * N_T + n]);
          }
        }
      }
      __syncthreads();
    }
    for (int e = tid; e < m * N_T; e += 256) {
      int r = e / N_T, n = e - r * N_T;
      int bl = n / D, d = n - bl * D;
      if (n < n_used && b0 + bl < B) dx0[((b0 + bl) * m + r) * (int64_t)D + d] += dx0s[e];
    }
  }
}

template <int N_T>
static size_t cin_dx_smem(int m) {
  return sizeof(float) * ((size_t)2 * m * N_T + CK * DX_KC + CK * N_T + (size_t)DX_KC * N_T);
}

// dxk must be zero-initialised by the caller is NOT required: this entry point clears it; dx0 is accumulated into (+=).
extern "C" int xdfm_cin_bwd_f32(const float* x0, const float* xk, int64_t xk_bstride, const float* W, const float* dy, int64_t B, int m,
                                int Hp, int H, int D, float* dW, float* db, float* dxk, float* dx0, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  XDFM_CHECK_ARG(D >= 1 && D <= 128, "cin_bwd_f32: D=%d unsupported (1..128)", D);
  if (B == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int K = Hp * m;
  XDFM_CHECK_ARG(workspace_bytes >= xdfm_cin_bwd_f32_workspace_bytes(B, m, Hp, H, D), "cin_bwd_f32: workspace too small");
  int S = cin_dw_splits(B, H, K);
  float* part = (float*)workspace;
  float* dbpart = part + (int64_t)S * H * K;
  if (dW != nullptr) {
    int64_t sps = ceil_div64(B, S);
    dim3 grid((unsigned)ceil_div64(K, 64), (unsigned)ceil_div64(H, 64), (unsigned)S);
    cin_dw_f32_kernel<<<grid, 256, 0, st>>>(x0, xk, xk_bstride, dy, B, m, Hp, H, D, sps, part);
    XDFM_LAUNCH_CHECK();
    int blocks = (int)min((int64_t)xdfm_num_sms() * 4, ceil_div64((int64_t)H * K, 256));
    cin_dw_reduce_kernel<<<blocks, 256, 0, st>>>(part, S, (int64_t)H * K, dW);
    XDFM_LAUNCH_CHECK();
  }
  if (db != nullptr) {
    dim3 g1((unsigned)ceil_div64(H, 64), DB_CHUNKS);
    cin_db_stage1<<<g1, 64, 0, st>>>(dy, B, H, D, dbpart);
    XDFM_LAUNCH_CHECK();
    cin_db_stage2<<<(unsigned)ceil_div64(H, 64), 64, 0, st>>>(dbpart, H, db);
    XDFM_LAUNCH_CHECK();
  }
  if (dxk != nullptr || dx0 != nullptr) {
    XDFM_CHECK_ARG(dxk != nullptr && dx0 != nullptr, "cin_bwd_f32: dxk and dx0 must both be given");
    XDFM_CUDA(cudaMemsetAsync(dxk, 0, (size_t)B * Hp * D * sizeof(float), st));
#define CIN_DX_LAUNCH(NT)                                                                                              \
  {                                                                                                                    \
    size_t sm = cin_dx_smem<NT>(m);                                                                                    \
    XDFM_CUDA(cudaFuncSetAttribute(cin_dx_f32_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
    int TB = NT / D;                                                                                                   \
    int64_t tiles = ceil_div64(B, TB);                                                                                 \
    int blocks = (int)std::min<int64_t>(tiles, (int64_t)xdfm_num_sms() * 2);                                                \
    cin_dx_f32_kernel<NT><<<blocks, 256, sm, st>>>(x0, xk, xk_bstride, W, dy, B, m, Hp, H, D, dxk, dx0);               \
  }
    if (cin_dx_smem<128>(m) <= 200 * 1024) CIN_DX_LAUNCH(128)
    else {
      xdfm_set_error("cin_bwd_f32: m=%d does not fit shared memory", m);
      return XDFM_ERR_UNSUPPORTED;
    }
#undef CIN_DX_LAUNCH
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}
