// Narrow dense layers over very tall activations: y_q[R, N] = act(x[R, K] . W_q[N, K]^T + b), K, N <= 32, q < 3.
//
// Replaces (reference, file:line): the bias-free E x E projections W_q / W_k / W_v / W_o of MultiHeadSelfAttention
// (deepctr/layers/cin_attention.py:49-61, 73-79, 95) and the Linear(E, E) -> Tanh -> Linear(E, 1) score MLP of AttentionPooling
// (deepctr/layers/cin_attention.py:114-118, 135-136) plus torch autograd of those nn.Linear modules.  At BASELINE config 3 these
// run over R = B * L = 16384 * 256 = 4.2 M rows of E = 16 floats: a GEMM with K = N = 16 is pure HBM streaming (64 B in, 64 B out
// per row and projection), so the tensor-core GEMM path (bf16 conversion passes + 128-row tiles) only adds traffic.  Here:
//   forward   one thread per row, the row in registers, the (up to three) weight matrices in shared memory (128-bit broadcast
//             reads): Q, K and V come out of ONE pass over x.
//   dX        the same kernel with the roles swapped: dx[R, K] = sum_q dy_q[R, N] . W_q[N, K].
//   dW, db    4 x 4 register blocks per thread over 128-row shared-memory tiles, rows dealt round-robin to thread groups, the
//             groups / blocks combined in a fixed order (two stages): bit-reproducible.
#include "common.cuh"
#include "../../include/xdfm.h"

#define SL_THREADS 256
#define SL_TILE 128
#define SL_MAXQ 3

__device__ __forceinline__ float sl_act(float v, int act) {
  if (act == XDFM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == XDFM_ACT_TANH) return tanhf(v);
  if (act == XDFM_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

struct SlRowParams {
  const float* in[SL_MAXQ];     // NIN inputs, each [R, I]
  const float* w[SL_MAXQ];      // weight matrices [N, K] row-major (forward: one per output; dX: one per input)
  float* out[SL_MAXQ];          // nq outputs, each [R, O]
  const float* bias;            // [O] or null (forward, nq == 1 only)
  int64_t R;
  int I, O, nq, act, transpose; // transpose = 0: forward (I = K, O = N); 1: dX (I = N, O = K, weights read transposed)
};

// IM = I rounded up to 8/16/32 (compile time: the input rows live in registers), NIN = number of concatenated inputs, RPT = rows
// per thread (every 128-bit weight load from shared memory applied to RPT rows; see sl_launch_rows for what was measured).
template <int IM, int NIN, int RPT>
__global__ void __launch_bounds__(SL_THREADS) sl_rows_kernel(SlRowParams p) {
  extern __shared__ __align__(16) float sl_smem[];
  constexpr int TL = IM * NIN;                       // length of the concatenated (padded) input vector
  const int OM = (p.O + 3) & ~3;
  // Wsm[q][o][j * IM + i]: zero padded
  for (int e = threadIdx.x; e < p.nq * OM * TL; e += SL_THREADS) {
    const int i = e % IM, j = (e / IM) % NIN, o = (e / TL) % OM, q = e / (TL * OM);
    float v = 0.f;
    if (i < p.I && o < p.O) v = p.transpose ? p.w[j][(size_t)i * p.O + o] : p.w[q][(size_t)o * p.I + i];
    sl_smem[e] = v;
  }
  __syncthreads();
  const bool vec_in = (p.I % 4) == 0, vec_out = (p.O % 4) == 0;
  for (int64_t r0 = ((int64_t)blockIdx.x * SL_THREADS + threadIdx.x) * RPT; r0 < p.R; r0 += (int64_t)gridDim.x * SL_THREADS * RPT) {
    float in[RPT][TL];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const bool live = r0 + u < p.R;
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const float* src = p.in[j] + (r0 + u) * p.I;
        if (vec_in) {
#pragma unroll
          for (int i4 = 0; i4 < IM / 4; ++i4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && i4 * 4 < p.I) v = __ldg(reinterpret_cast<const float4*>(src) + i4);
            in[u][j * IM + i4 * 4 + 0] = v.x; in[u][j * IM + i4 * 4 + 1] = v.y;
            in[u][j * IM + i4 * 4 + 2] = v.z; in[u][j * IM + i4 * 4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < IM; ++i) in[u][j * IM + i] = (live && i < p.I) ? __ldg(src + i) : 0.f;
        }
      }
    }
    for (int q = 0; q < p.nq; ++q) {
      for (int o0 = 0; o0 < p.O; o0 += 4) {
        float a[RPT][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4* w = reinterpret_cast<const float4*>(sl_smem + ((size_t)q * OM + o0 + t) * TL);
          float s0[RPT], s1[RPT];
#pragma unroll
          for (int u = 0; u < RPT; ++u) { s0[u] = 0.f; s1[u] = 0.f; }
#pragma unroll
          for (int i4 = 0; i4 < TL / 4; i4 += 2) {
            const float4 wa = w[i4], wb = w[i4 + 1];
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
              s0[u] = fmaf(in[u][i4 * 4 + 0], wa.x, s0[u]); s0[u] = fmaf(in[u][i4 * 4 + 1], wa.y, s0[u]);
              s0[u] = fmaf(in[u][i4 * 4 + 2], wa.z, s0[u]); s0[u] = fmaf(in[u][i4 * 4 + 3], wa.w, s0[u]);
              s1[u] = fmaf(in[u][i4 * 4 + 4], wb.x, s1[u]); s1[u] = fmaf(in[u][i4 * 4 + 5], wb.y, s1[u]);
              s1[u] = fmaf(in[u][i4 * 4 + 6], wb.z, s1[u]); s1[u] = fmaf(in[u][i4 * 4 + 7], wb.w, s1[u]);
            }
          }
          const float bv = (p.bias != nullptr && o0 + t < p.O) ? __ldg(p.bias + o0 + t) : 0.f;
#pragma unroll
          for (int u = 0; u < RPT; ++u) a[u][t] = sl_act(s0[u] + s1[u] + bv, p.act);
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
          if (r0 + u < p.R) {
            float* dst = p.out[q] + (r0 + u) * p.O;
            if (vec_out) {
              *reinterpret_cast<float4*>(dst + o0) = make_float4(a[u][0], a[u][1], a[u][2], a[u][3]);
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (o0 + t < p.O) dst[o0 + t] = a[u][t];
            }
          }
        }
      }
    }
  }
}

// Coalesced variant for I == IM (16 or 32 floats per row, a multiple of 4) and O % 4 == 0: a warp owns 32 consecutive rows = one
// contiguous block of memory; it moves the block with lane-contiguous 128-bit accesses through a private shared-memory patch
// ([32][IM + 4] floats: the padded pitch makes both the block-order stores and the row-order loads conflict-free) instead of
// having every lane walk its own 64-byte row (32 half-used sectors per instruction).
template <int IM, int NIN, int RPT>
__global__ void __launch_bounds__(SL_THREADS) sl_rows_staged_kernel(SlRowParams p) {
  extern __shared__ __align__(16) float sl_smem[];
  constexpr int TL = IM * NIN;
  constexpr int PITCH = IM + 4;
  constexpr int ROWS = 32 * RPT;                      // rows a warp moves and computes per iteration (thread: rows lane + 32 u)
  const int OM = p.O;                                 // multiple of 4
  const int n_w = p.nq * OM * TL;
  for (int e = threadIdx.x; e < n_w; e += SL_THREADS) {
    const int i = e % IM, j = (e / IM) % NIN, o = (e / TL) % OM, q = e / (TL * OM);
    sl_smem[e] = p.transpose ? p.w[j][(size_t)i * p.O + o] : p.w[q][(size_t)o * p.I + i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* patch = sl_smem + ((n_w + 3) & ~3) + warp * ROWS * PITCH;
  constexpr int IQ = IM / 4;                          // float4s per input row
  const int OQ = p.O / 4;                             // float4s per output row
  for (int64_t rb = ((int64_t)blockIdx.x * (SL_THREADS / 32) + warp) * ROWS; rb < p.R; rb += (int64_t)gridDim.x * (SL_THREADS / 32) * ROWS) {
    float in[RPT][TL];
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      const float4* src = reinterpret_cast<const float4*>(p.in[j]) + rb * IQ;
#pragma unroll
      for (int c = 0; c < IQ * RPT; ++c) {
        const int idx = c * 32 + lane, row = idx / IQ, part = idx % IQ;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rb + row < p.R) v = ldg_nc_f4(src + idx);
        *reinterpret_cast<float4*>(patch + row * PITCH + part * 4) = v;
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < RPT; ++u)
#pragma unroll
        for (int i4 = 0; i4 < IQ; ++i4) {
          const float4 v = *reinterpret_cast<const float4*>(patch + (lane + 32 * u) * PITCH + i4 * 4);
          in[u][j * IM + i4 * 4 + 0] = v.x; in[u][j * IM + i4 * 4 + 1] = v.y;
          in[u][j * IM + i4 * 4 + 2] = v.z; in[u][j * IM + i4 * 4 + 3] = v.w;
        }
      __syncwarp();
    }
    for (int q = 0; q < p.nq; ++q) {
      for (int o0 = 0; o0 < p.O; o0 += 4) {
        float a[RPT][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4* w = reinterpret_cast<const float4*>(sl_smem + ((size_t)q * OM + o0 + t) * TL);
          float s0[RPT], s1[RPT];
#pragma unroll
          for (int u = 0; u < RPT; ++u) { s0[u] = 0.f; s1[u] = 0.f; }
#pragma unroll
          for (int i4 = 0; i4 < TL / 4; i4 += 2) {
            const float4 wa = w[i4], wb = w[i4 + 1];      // one weight load serves RPT rows
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
              s0[u] = fmaf(in[u][i4 * 4 + 0], wa.x, s0[u]); s0[u] = fmaf(in[u][i4 * 4 + 1], wa.y, s0[u]);
              s0[u] = fmaf(in[u][i4 * 4 + 2], wa.z, s0[u]); s0[u] = fmaf(in[u][i4 * 4 + 3], wa.w, s0[u]);
              s1[u] = fmaf(in[u][i4 * 4 + 4], wb.x, s1[u]); s1[u] = fmaf(in[u][i4 * 4 + 5], wb.y, s1[u]);
              s1[u] = fmaf(in[u][i4 * 4 + 6], wb.z, s1[u]); s1[u] = fmaf(in[u][i4 * 4 + 7], wb.w, s1[u]);
            }
          }
          const float bv = p.bias != nullptr ? __ldg(p.bias + o0 + t) : 0.f;
#pragma unroll
          for (int u = 0; u < RPT; ++u) a[u][t] = sl_act(s0[u] + s1[u] + bv, p.act);
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          *reinterpret_cast<float4*>(patch + (lane + 32 * u) * PITCH + o0) = make_float4(a[u][0], a[u][1], a[u][2], a[u][3]);
      }
      __syncwarp();
      float4* dst = reinterpret_cast<float4*>(p.out[q]) + rb * OQ;
      for (int c = 0; c < OQ * RPT; ++c) {
        const int idx = c * 32 + lane, row = idx / OQ, part = idx % OQ;
        if (rb + row < p.R) dst[idx] = *reinterpret_cast<const float4*>(patch + row * PITCH + part * 4);
      }
      __syncwarp();
    }
  }
}

// rows per thread of the coalesced kernel (0 = per-lane-row kernel; >= 8: also for concatenated inputs, diagnostic).  Measured at
// K = N = 16, R = 4.2 M: one input, one / three outputs 0.163 / 0.385 ms (0), 0.163 / 0.386 (1), 0.147 / 0.344 (2), 0.163 / 0.385 (4);
// three inputs (dX of Q/K/V) 0.348 (0), 0.413 (1), 0.574 (2) -> 2 for single-input launches, the per-lane kernel otherwise
int g_sl_staged = 2;
extern "C" void xdfm_small_linear_set_staged(int v) { g_sl_staged = v < 0 ? 0 : v; }

static int sl_check(int64_t R, int K, int N, int nq, const char* what) {
  XDFM_CHECK_ARG(R >= 0 && K >= 1 && N >= 1 && nq >= 1 && nq <= SL_MAXQ, "%s: bad shape R=%lld K=%d N=%d n=%d", what, (long long)R, K, N, nq);
  if (K > 32 || N > 32) {
    xdfm_set_error("%s: K=%d, N=%d: the narrow-layer kernels cover K, N <= 32 (use xdfm_gemm_f32 / xdfm_gemm_tc)", what, K, N);
    return XDFM_ERR_UNSUPPORTED;
  }
  return XDFM_OK;
}

template <int IM>
static int sl_launch_rows(const SlRowParams& p, int nin, cudaStream_t st) {
  if (g_sl_staged && (nin == 1 || g_sl_staged >= 8) && p.I == IM && (p.O & 3) == 0 && p.O <= IM) {
    // g_sl_staged = rows per thread (1, 2 or 4) of the coalesced kernel; three concatenated inputs keep 1 or 2 (registers)
    constexpr int RA = IM <= 16 ? 4 : 2;
    int rpt = g_sl_staged >= 4 ? RA : (g_sl_staged >= 2 ? 2 : 1);
    if (nin > 1) rpt = std::min(rpt, IM <= 16 ? 2 : 1);
    const size_t smem = ((((size_t)p.nq * p.O * IM * nin + 3) & ~(size_t)3) + (size_t)(SL_THREADS / 32) * 32 * rpt * (IM + 4)) * sizeof(float);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(p.R, (int64_t)SL_THREADS * rpt), (int64_t)xdfm_num_sms() * 8));
#define SL_STAGED(NI, RP)                                                                                              \
    do {                                                                                                               \
      XDFM_CUDA(cudaFuncSetAttribute(sl_rows_staged_kernel<IM, NI, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      sl_rows_staged_kernel<IM, NI, RP><<<blocks, SL_THREADS, smem, st>>>(p);                                          \
    } while (0)
    if (nin == 1) { if (rpt == RA) SL_STAGED(1, RA); else if (rpt == 2) SL_STAGED(1, 2); else SL_STAGED(1, 1); }
    else if (nin == 2) { if (rpt == 2) SL_STAGED(2, 2); else SL_STAGED(2, 1); }
    else { if (rpt == 2) SL_STAGED(3, 2); else SL_STAGED(3, 1); }
#undef SL_STAGED
    XDFM_LAUNCH_CHECK();
    return XDFM_OK;
  }
  const int OM = (p.O + 3) & ~3;
  const size_t smem = (size_t)p.nq * OM * IM * nin * sizeof(float);
  // rows per thread (every weight load applied to RPT rows): measured SLOWER at BASELINE config 3 (K = N = 16: 0.198 vs 0.172 ms
  // for one output, 0.572 vs 0.381 ms for three) -- the shared-memory weight loads are not what bounds this kernel; kept at 1
  constexpr int RPT1 = 1, RPT2 = 1, RPT3 = 1;        // (2 rows per thread: 0.211 ms; 4: 0.198 ms; 1: 0.172 ms)
  const int rpt = nin == 1 ? RPT1 : (nin == 2 ? RPT2 : RPT3);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(p.R, (int64_t)SL_THREADS * rpt), (int64_t)xdfm_num_sms() * 8));
  if (nin == 1) sl_rows_kernel<IM, 1, RPT1><<<blocks, SL_THREADS, smem, st>>>(p);
  else if (nin == 2) sl_rows_kernel<IM, 2, RPT2><<<blocks, SL_THREADS, smem, st>>>(p);
  else sl_rows_kernel<IM, 3, RPT3><<<blocks, SL_THREADS, smem, st>>>(p);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

static int sl_dispatch_rows(const SlRowParams& p, int nin, cudaStream_t st) {
  if (p.I <= 8) return sl_launch_rows<8>(p, nin, st);
  if (p.I <= 16) return sl_launch_rows<16>(p, nin, st);
  return sl_launch_rows<32>(p, nin, st);
}

extern "C" int xdfm_small_linear_fwd(const float* x, const float* w0, const float* w1, const float* w2, const float* bias, int act,
                                     int64_t R, int K, int N, int nout, float* y0, float* y1, float* y2, void* stream) {
  int rc = sl_check(R, K, N, nout, "small_linear_fwd");
  if (rc) return rc;
  XDFM_CHECK_ARG(bias == nullptr || nout == 1, "small_linear_fwd: a bias needs nout == 1");
  XDFM_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y0 % 16 == 0) && ((uintptr_t)y1 % 16 == 0) && ((uintptr_t)y2 % 16 == 0),
                 "small_linear_fwd: x / y must be 16-byte aligned");
  if (R == 0) return XDFM_OK;
  SlRowParams p = {};
  p.in[0] = x; p.w[0] = w0; p.w[1] = w1; p.w[2] = w2; p.out[0] = y0; p.out[1] = y1; p.out[2] = y2;
  p.bias = bias; p.R = R; p.I = K; p.O = N; p.nq = nout; p.act = act; p.transpose = 0;
  return sl_dispatch_rows(p, 1, (cudaStream_t)stream);
}

extern "C" int xdfm_small_linear_bwd_dx(const float* dy0, const float* dy1, const float* dy2, const float* w0, const float* w1,
                                        const float* w2, int64_t R, int K, int N, int nout, float* dx, void* stream) {
  int rc = sl_check(R, K, N, nout, "small_linear_bwd_dx");
  if (rc) return rc;
  XDFM_CHECK_ARG(((uintptr_t)dx % 16 == 0) && ((uintptr_t)dy0 % 16 == 0) && ((uintptr_t)dy1 % 16 == 0) && ((uintptr_t)dy2 % 16 == 0),
                 "small_linear_bwd_dx: dx / dy must be 16-byte aligned");
  if (R == 0) return XDFM_OK;
  SlRowParams p = {};
  p.in[0] = dy0; p.in[1] = dy1; p.in[2] = dy2; p.w[0] = w0; p.w[1] = w1; p.w[2] = w2; p.out[0] = dx;
  p.bias = nullptr; p.R = R; p.I = N; p.O = K; p.nq = 1; p.act = XDFM_ACT_NONE; p.transpose = 1;
  return sl_dispatch_rows(p, nout, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// dW_q[n, k] = sum_r dy_q[r, n] x[r, k];  db[n] = sum_r dy_0[r, n]
// ------------------------------------------------------------------------------------------------
struct SlDwParams {
  const float* x;               // [R, K]
  const float* dy[SL_MAXQ];     // [R, N]
  float* partial;               // [blocks, nq * N * K + N]
  int64_t R;
  int K, N, nq, KM, NM, KB, NB, TG, G, want_db;
};

// thread (g, t): group g = tid / TG handles rows g, g + G, ... of every tile; t -> 4 x 4 block (n0 = (t / KB) * 4, k0 = (t % KB) * 4)
__global__ void __launch_bounds__(SL_THREADS) sl_dw_kernel(SlDwParams p) {
  extern __shared__ __align__(16) float sl_smem[];
  float* sx = sl_smem;                                   // [TILE][KM]
  float* sdy = sx + SL_TILE * p.KM;                      // [nq][TILE][NM]
  const int tid = threadIdx.x;
  const int g = tid / p.TG, t = tid % p.TG;
  const bool active = g < p.G;
  const int n0 = (t / p.KB) * 4, k0 = (t % p.KB) * 4;
  float acc[SL_MAXQ][16];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < SL_MAXQ; ++q)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[q][i] = 0.f;
  const bool vec_x = (p.K == p.KM), vec_y = (p.N == p.NM);
  for (int64_t r0 = (int64_t)blockIdx.x * SL_TILE; r0 < p.R; r0 += (int64_t)gridDim.x * SL_TILE) {
    const int rows = (int)min((int64_t)SL_TILE, p.R - r0);
    __syncthreads();                                     // previous tile consumed
    if (vec_x) {
      const float4* src = reinterpret_cast<const float4*>(p.x + r0 * p.K);
      for (int e = tid; e < SL_TILE * p.KM / 4; e += SL_THREADS)
        reinterpret_cast<float4*>(sx)[e] = e < rows * p.KM / 4 ? ldg_nc_f4(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int e = tid; e < SL_TILE * p.KM; e += SL_THREADS) {
        const int r = e / p.KM, k = e % p.KM;
        sx[e] = (r < rows && k < p.K) ? __ldg(p.x + (r0 + r) * p.K + k) : 0.f;
      }
    }
    for (int q = 0; q < p.nq; ++q) {
      float* d = sdy + (size_t)q * SL_TILE * p.NM;
      if (vec_y) {
        const float4* src = reinterpret_cast<const float4*>(p.dy[q] + r0 * p.N);
        for (int e = tid; e < SL_TILE * p.NM / 4; e += SL_THREADS)
          reinterpret_cast<float4*>(d)[e] = e < rows * p.NM / 4 ? ldg_nc_f4(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int e = tid; e < SL_TILE * p.NM; e += SL_THREADS) {
          const int r = e / p.NM, n = e % p.NM;
          d[e] = (r < rows && n < p.N) ? __ldg(p.dy[q] + (r0 + r) * p.N + n) : 0.f;
        }
      }
    }
    __syncthreads();
    if (active) {
      for (int r = g; r < SL_TILE; r += p.G) {           // rows past `rows` are zero: fixed trip count, fixed order
        const float4 xv = *reinterpret_cast<const float4*>(sx + r * p.KM + k0);
#pragma unroll
        for (int q = 0; q < SL_MAXQ; ++q) {
          if (q < p.nq) {
            const float4 dv = *reinterpret_cast<const float4*>(sdy + ((size_t)q * SL_TILE + r) * p.NM + n0);
            const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              acc[q][a * 4 + 0] = fmaf(d4[a], xv.x, acc[q][a * 4 + 0]);
              acc[q][a * 4 + 1] = fmaf(d4[a], xv.y, acc[q][a * 4 + 1]);
              acc[q][a * 4 + 2] = fmaf(d4[a], xv.z, acc[q][a * 4 + 2]);
              acc[q][a * 4 + 3] = fmaf(d4[a], xv.w, acc[q][a * 4 + 3]);
            }
            if (q == 0 && k0 == 0) { accb[0] += dv.x; accb[1] += dv.y; accb[2] += dv.z; accb[3] += dv.w; }
          }
        }
      }
    }
  }
  // ---- combine the groups in ascending order: red[g][q][n][k] (padded NM x KM) + redb[g][NM]
  __syncthreads();
  float* red = sl_smem;
  const int per_g = p.nq * p.NM * p.KM + p.NM;
  if (active) {
    float* mine = red + (size_t)g * per_g;
#pragma unroll
    for (int q = 0; q < SL_MAXQ; ++q)
      if (q < p.nq)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) mine[((size_t)q * p.NM + n0 + a) * p.KM + k0 + b] = acc[q][a * 4 + b];
    if (k0 == 0)
#pragma unroll
      for (int a = 0; a < 4; ++a) mine[p.nq * p.NM * p.KM + n0 + a] = accb[a];
  }
  __syncthreads();
  const int n_out = p.nq * p.N * p.K + (p.want_db ? p.N : 0);
  float* dst = p.partial + (size_t)blockIdx.x * (p.nq * p.N * p.K + p.N);
  for (int e = tid; e < n_out; e += SL_THREADS) {
    int src;
    if (e < p.nq * p.N * p.K) {
      const int k = e % p.K, n = (e / p.K) % p.N, q = e / (p.K * p.N);
      src = (q * p.NM + n) * p.KM + k;
    } else {
      src = p.nq * p.NM * p.KM + (e - p.nq * p.N * p.K);
    }
    float s = 0.f;
    for (int gg = 0; gg < p.G; ++gg) s += red[(size_t)gg * per_g + src];
    dst[e] = s;
  }
}

// one warp per output element: lane l adds the partials of blocks l, l + 32, ... in ascending order, then a fixed xor-shuffle tree
__global__ void __launch_bounds__(256) sl_dw_reduce_kernel(const float* __restrict__ partial, int blocks, int stride, int nw, int N,
                                                           float* dw0, float* dw1, float* dw2, int per_w, float* db) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (e >= nw + (db != nullptr ? N : 0)) return;
  float s = 0.f;
  for (int b = lane; b < blocks; b += 32) s += partial[(size_t)b * stride + e];
  s = warp_sum(s);
  if (lane == 0) {
    if (e < nw) {
      const int q = e / per_w, i = e % per_w;
      (q == 0 ? dw0 : (q == 1 ? dw1 : dw2))[i] = s;
    } else {
      db[e - nw] = s;
    }
  }
}

static void sl_dw_geom(int64_t R, int K, int N, int nq, SlDwParams* p, int* blocks, size_t* smem) {
  p->K = K; p->N = N; p->nq = nq;
  p->KM = (K + 3) & ~3; p->NM = (N + 3) & ~3;
  p->KB = p->KM / 4; p->NB = p->NM / 4;
  p->TG = p->KB * p->NB;
  p->G = std::min(SL_THREADS / p->TG, SL_TILE);
  const size_t tiles = (size_t)SL_TILE * (p->KM + (size_t)nq * p->NM);
  const size_t red = (size_t)p->G * ((size_t)nq * p->NM * p->KM + p->NM);
  *smem = std::max(tiles, red) * sizeof(float);
  // a block alternates load -> sync -> a few hundred FMAs -> sync per 128-row tile: it is latency-bound on its own, so the SM is
  // filled with as many blocks as the tiles' shared memory allows (round 1: 2 blocks per SM ran at 1.2 TB/s)
  *blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(R, SL_TILE), (int64_t)xdfm_num_sms() * 8));
}

extern "C" int64_t xdfm_small_linear_bwd_dw_workspace_bytes(int64_t R, int K, int N, int nout) {
  if (K < 1 || N < 1 || K > 32 || N > 32 || nout < 1 || nout > SL_MAXQ) return -1;
  SlDwParams p;
  int blocks;
  size_t smem;
  sl_dw_geom(std::max<int64_t>(R, 1), K, N, nout, &p, &blocks, &smem);
  return (int64_t)blocks * ((int64_t)nout * N * K + N) * (int64_t)sizeof(float);
}

extern "C" int xdfm_small_linear_bwd_dw(const float* x, const float* dy0, const float* dy1, const float* dy2, int64_t R, int K, int N,
                                        int nout, float* dw0, float* dw1, float* dw2, float* db, void* workspace, void* stream) {
  int rc = sl_check(R, K, N, nout, "small_linear_bwd_dw");
  if (rc) return rc;
  XDFM_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)dy0 % 16 == 0) && ((uintptr_t)dy1 % 16 == 0) && ((uintptr_t)dy2 % 16 == 0),
                 "small_linear_bwd_dw: x / dy must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    float* dws[3] = {dw0, dw1, dw2};
    for (int q = 0; q < nout; ++q) XDFM_CUDA(cudaMemsetAsync(dws[q], 0, (size_t)N * K * sizeof(float), st));
    if (db != nullptr) XDFM_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
    return XDFM_OK;
  }
  XDFM_CHECK_ARG(workspace != nullptr, "small_linear_bwd_dw: workspace is null");
  SlDwParams p = {};
  int blocks;
  size_t smem;
  sl_dw_geom(R, K, N, nout, &p, &blocks, &smem);
  p.x = x; p.dy[0] = dy0; p.dy[1] = dy1; p.dy[2] = dy2; p.partial = (float*)workspace; p.R = R; p.want_db = db != nullptr;
  XDFM_CUDA(cudaFuncSetAttribute(sl_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sl_dw_kernel<<<blocks, SL_THREADS, smem, st>>>(p);
  XDFM_LAUNCH_CHECK();
  const int nw = nout * N * K;
  sl_dw_reduce_kernel<<<(nw + N + 7) / 8, 256, 0, st>>>((const float*)workspace, blocks, nw + N, nw, N, dw0, dw1, dw2, N * K, db);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
