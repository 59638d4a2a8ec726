// Dense layers on the tensor cores: C[M,N] = act(A[M,K] . B[N,K]^T + bias[N]), bf16 operands, fp32 accumulate in TMEM.
//
// Replaces (reference, file:line): nn.Linear + ReLU of DNN.forward (deepctr/layers/core.py:120-134) and its autograd
// (dX = dY.W, dW = dY^T.X) when the model runs in bf16 precision (BASELINE config 2 "bf16"); the fp32 SGEMM of dense.cu stays
// the exact-precision mode.
//
// Both operands are K-major bf16 matrices (row pitch a multiple of 8 elements) produced by xdfm_cvt_bf16 -- which also does the
// transposes the backward GEMMs need -- and streamed by TMA (SWIZZLE_128B, 64-wide K chunks, OOB rows/cols zero-filled) through
// a 4-stage mbarrier ring; one elected thread issues tcgen05.mma kind::f16 (M=128, N=BN<=256, K=16) with both operands in
// shared memory; four epilogue warps read the accumulator back with tcgen05.ld and apply bias + activation.  One CTA per
// 128 x BN output tile and K split; split-K partials are reduced in a fixed order by a second kernel (deterministic).
#include "tc_common.cuh"
#include "../../include/xdfm.h"

using namespace tc;

#define GT_THREADS 192     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
#define GT_STAGES 4

struct GemmTcParams {
  float* C;                 // [M, ldc] (splits == 1)
  float* partial;           // [splits, M, N] (splits > 1)
  const float* bias;
  int M, N, K, ldc, BN, act, splits, chunks_total, chunks_per_split;
};

struct __align__(8) GemmTcBars {
  uint64_t full[GT_STAGES], empty[GT_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == XDFM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == XDFM_ACT_TANH) return tanhf(v);
  if (act == XDFM_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)p.BN * 128;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  GemmTcBars* bars = reinterpret_cast<GemmTcBars*>(smem + (size_t)GT_STAGES * stage_bytes);
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.BN, split = blockIdx.z;
  const int c_begin = split * p.chunks_per_split;
  const int c_end = min(p.chunks_total, c_begin + p.chunks_per_split);
  const int nchunks = c_end - c_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmB);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % GT_STAGES;
        if (c >= GT_STAGES) mbar_wait(&bars->empty[s], ((c / GT_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&bars->full[s], a_bytes + b_bytes);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        tma_load_2d(sa, &tmA, (c_begin + c) * 64, m0, &bars->full[s]);
        tma_load_2d(sa + a_bytes, &tmB, (c_begin + c) * 64, n0, &bars->full[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.BN);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % GT_STAGES;
      mbar_wait(&bars->full[s], (c / GT_STAGES) & 1);
      fence_after_sync();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t adesc = make_desc_k_sw128(sa), bdesc = make_desc_k_sw128(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars->empty[s]);
        if (c == nchunks - 1) umma_commit(&bars->acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue (4 warps): TMEM -> (+bias, activation) -> shared-memory tile -> global memory, whole rows at a time.
    // A thread owns one accumulator ROW; writing it straight to global memory touches 32 different rows per store instruction
    // (one 4-byte piece of 32 sectors each).  The operand ring is free once the accumulator is complete, so the tile is staged
    // there (pitch BN + 4 floats: conflict-free 128-bit stores) and every warp then streams complete rows, lane-contiguous.
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int rl = q * 32 + lane;                 // row of the tile owned by this thread
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (nchunks > 0) {
      mbar_wait(&bars->acc_full, 0);              // every MMA has completed: the ring is no longer read, all TMA loads have landed
      fence_after_sync();
    }
    float* stile = reinterpret_cast<float*>(smem);
    const int pitch = p.BN + 4;
    const bool fin = p.splits == 1;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      if (nchunks > 0) {
        tmem_ld_x16(taddr + c0, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      float y[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = n0 + c0 + i;
        float t = __uint_as_float(v[i]);
        if (fin) {
          if (p.bias != nullptr && col < p.N) t += __ldg(p.bias + col);
          t = apply_act(t, p.act);
        }
        y[i] = t;
      }
      float4* d4 = reinterpret_cast<float4*>(stile + (size_t)rl * pitch + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) d4[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");         // the four epilogue warps: tile complete
    const int ncols = min(p.BN, p.N - n0);                  // valid columns of this tile
    const int64_t ld = p.splits > 1 ? p.N : p.ldc;
    float* base = p.splits > 1 ? p.partial + (size_t)split * p.M * p.N : p.C;
    const bool vec = ((ld & 3) == 0) && ((n0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    for (int rr = q; rr < 128; rr += 4) {
      const int row = m0 + rr;
      if (row >= p.M) break;
      const float* srow = stile + (size_t)rr * pitch;
      float* drow = base + (size_t)row * ld + n0;
      if (vec) {
        for (int c4 = lane; c4 * 4 < ncols; c4 += 32) {
          const float4 t = *reinterpret_cast<const float4*>(srow + c4 * 4);
          if (c4 * 4 + 4 <= ncols) {
            *reinterpret_cast<float4*>(drow + c4 * 4) = t;
          } else {
            const float e[4] = {t.x, t.y, t.z, t.w};
            for (int i = 0; c4 * 4 + i < ncols; ++i) drow[c4 * 4 + i] = e[i];
          }
        }
      } else {
        for (int c = lane; c < ncols; c += 32) drow[c] = srow[c];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

__global__ void gemm_tc_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, const float* __restrict__ bias, int act,
                                      float* __restrict__ C, int ldc) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * total + e];
    const int col = (int)(e % N);
    if (bias != nullptr) s += bias[col];
    C[(e / N) * ldc + col] = apply_act(s, act);
  }
}

static int gemm_tc_plan(int M, int N, int K, int* BN, int* n_tiles, int* splits, int* chunks_per_split) {
  *n_tiles = (N + 255) / 256;
  *BN = ((N + *n_tiles - 1) / *n_tiles + 15) / 16 * 16;
  const int chunks = (K + 63) / 64;
  const int tiles = ((M + 127) / 128) * *n_tiles;
  int s = 1;
  const int sms = xdfm_num_sms();
  if (tiles * 2 <= sms) s = std::min(chunks, std::max(1, sms / tiles));     // few output tiles, long K: split the reduction
  *chunks_per_split = (chunks + s - 1) / s;
  *splits = (chunks + *chunks_per_split - 1) / *chunks_per_split;
  return chunks;
}

extern "C" int64_t xdfm_gemm_tc_workspace_bytes(int M, int N, int K) {
  int BN, nt, splits, cps;
  gemm_tc_plan(M, N, K, &BN, &nt, &splits, &cps);
  return splits > 1 ? (int64_t)splits * M * N * 4 : 256;
}

// A [M, K] bf16 row pitch lda, Bm [N, K] bf16 row pitch ldb (elements, multiples of 8; 16-byte aligned bases); C fp32 [M, ldc]
extern "C" int xdfm_gemm_tc(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, float* C, int ldc,
                            const float* bias, int act, void* workspace, int64_t workspace_bytes, void* stream) {
  XDFM_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && ldc >= N, "gemm_tc: bad shape M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  XDFM_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K && ((uintptr_t)A % 16 == 0) && ((uintptr_t)Bm % 16 == 0),
                 "gemm_tc: operands must be 16-byte aligned with row pitches (%lld, %lld) multiples of 8 and >= K=%d", (long long)lda,
                 (long long)ldb, K);
  if (M == 0) return XDFM_OK;
  GemmTcParams p;
  int n_tiles;
  p.chunks_total = gemm_tc_plan(M, N, K, &p.BN, &n_tiles, &p.splits, &p.chunks_per_split);
  p.C = C; p.partial = (float*)workspace; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.act = act;
  if (p.splits > 1)
    XDFM_CHECK_ARG(workspace != nullptr && workspace_bytes >= (int64_t)p.splits * M * N * 4, "gemm_tc: workspace too small");
  CUtensorMap tmA, tmB;
  int rc = xdfm_make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 128, 64, 1);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmB, Bm, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, (uint32_t)p.BN, 64, 1);
  if (rc) return rc;
  const uint32_t stage_bytes = 128 * 128 + (((uint32_t)p.BN * 128 + 1023) & ~1023u);
  const size_t smem = (size_t)GT_STAGES * stage_bytes + sizeof(GemmTcBars) + 64;
  cudaStream_t st = (cudaStream_t)stream;
  XDFM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((M + 127) / 128, n_tiles, p.splits);
  gemm_tc_kernel<<<grid, GT_THREADS, smem, st>>>(tmA, tmB, p);
  XDFM_LAUNCH_CHECK();
  if (p.splits > 1) {
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64((int64_t)M * N, 256));
    gemm_tc_reduce_kernel<<<blocks, 256, 0, st>>>(p.partial, p.splits, M, N, bias, act, C, ldc);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 [R, C] (row pitch ld) -> bf16, either [R, CP] (transpose = 0) or [C, RP] (transpose = 1); padding columns are zero
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cvt_bf16_kernel(const float* __restrict__ src, int R, int C, int64_t ld, __nv_bfloat16* __restrict__ dst,
                                                       int64_t dpitch, int dcols) {
  const int64_t total = (int64_t)R * dcols;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / dcols;
    const int c = (int)(e - r * dcols);
    dst[r * dpitch + c] = __float2bfloat16(c < C ? src[r * ld + c] : 0.f);
  }
}

__global__ void __launch_bounds__(256) cvt_bf16_t_kernel(const float* __restrict__ src, int R, int C, int64_t ld,
                                                         __nv_bfloat16* __restrict__ dst, int64_t dpitch, int dcols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? src[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;          // dst row = c, dst col = r
    if (c < C && r < dcols) dst[(int64_t)c * dpitch + r] = __float2bfloat16(tile[tx][i]);
  }
}

extern "C" int xdfm_cvt_bf16(const float* src, int R, int C, int64_t ld, int transpose, void* dst, int64_t dst_pitch, void* stream) {
  const int drows = transpose ? C : R, dcols_true = transpose ? R : C;
  XDFM_CHECK_ARG(R >= 0 && C >= 1 && ld >= C && dst_pitch % 8 == 0 && dst_pitch >= dcols_true,
                 "cvt_bf16: bad shape R=%d C=%d ld=%lld dst_pitch=%lld", R, C, (long long)ld, (long long)dst_pitch);
  if (R == 0) return XDFM_OK;
  (void)drows;
  cudaStream_t st = (cudaStream_t)stream;
  const int dcols = (int)dst_pitch;            // write the padding too (zeros)
  if (!transpose) {
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64((int64_t)R * dcols, 256));
    cvt_bf16_kernel<<<blocks, 256, 0, st>>>(src, R, C, ld, (__nv_bfloat16*)dst, dst_pitch, dcols);
  } else {
    dim3 grid((dcols + 31) / 32, (C + 31) / 32);
    cvt_bf16_t_kernel<<<grid, 256, 0, st>>>(src, R, C, ld, (__nv_bfloat16*)dst, dst_pitch, dcols);
  }
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}


// ------------------------------------------------------------------------------------------------
// One pass that produces every bf16 operand a dense layer's tensor-core GEMMs need from an fp32 matrix:
//     g[r, c]   = src[r, c] * act'(y[r, c])          (y = NULL or act = NONE: g = src)
//     dst [R, dst_pitch]   = bf16(g)                  (row-major: A operand of  g . W)
//     dstT[C, dstT_pitch]  = bf16(g)^T                (A operand of  g^T . x)
//     colsum[c]            = sum_r g[r, c]            (fp32, fixed order: per 64-row tile, then over the tiles -- the bias gradient)
// Any of dst / dstT / colsum may be NULL.  Replaces act_bwd + cvt_bf16 + cvt_bf16(transpose) + wcolsum (four passes over the
// matrix, five launches) in the backward of a dense layer, and cvt_bf16 + cvt_bf16(transpose) in its forward.  64 x 64 tiles
// through shared memory: both outputs are written in 16- / 32-byte runs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cvt_bf16_both_kernel(const float* __restrict__ src, const float* __restrict__ y, int act, int R, int C,
                                                            int64_t ld, __nv_bfloat16* __restrict__ dst, int64_t dpitch,
                                                            __nv_bfloat16* __restrict__ dstT, int64_t tpitch, float* __restrict__ part) {
  __shared__ float tile[64][65];
  __shared__ float psum[4][64];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  {
    // 16 threads per row (one float4 each: 256 contiguous bytes), 16 rows per pass
    const int c4 = (threadIdx.x & 15) * 4;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                     (y == nullptr || (reinterpret_cast<uintptr_t>(y) & 15) == 0);
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int rr = (threadIdx.x >> 4) + pass * 16;
      const int r = r0 + rr, c = c0 + c4;
      float g[4] = {0.f, 0.f, 0.f, 0.f}, yv[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < R && c < C) {
        if (vec && c + 4 <= C) {
          const float4 t = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c);
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
          if (y != nullptr) {
            const float4 u = *reinterpret_cast<const float4*>(y + (int64_t)r * ld + c);
            yv[0] = u.x; yv[1] = u.y; yv[2] = u.z; yv[3] = u.w;
          }
        } else {
          for (int i = 0; i < 4; ++i)
            if (c + i < C) {
              g[i] = src[(int64_t)r * ld + c + i];
              if (y != nullptr) yv[i] = y[(int64_t)r * ld + c + i];
            }
        }
        if (y != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (act == XDFM_ACT_RELU) g[i] = yv[i] > 0.f ? g[i] : 0.f;
            else if (act == XDFM_ACT_TANH) g[i] = g[i] * (1.f - yv[i] * yv[i]);
            else if (act == XDFM_ACT_SIGMOID) g[i] = g[i] * yv[i] * (1.f - yv[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) tile[rr][c4 + i] = g[i];
      if (dst != nullptr && r < R && c < dpitch) {       // pitch is a multiple of 8: the four columns never straddle it
        const __nv_bfloat162 t0 = __floats2bfloat162_rn(g[0], g[1]), t1 = __floats2bfloat162_rn(g[2], g[3]);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&t0);
        o.y = *reinterpret_cast<const uint32_t*>(&t1);
        *reinterpret_cast<uint2*>(dst + (int64_t)r * dpitch + c) = o;
      }
    }
  }
  __syncthreads();
  {
    const int cc = threadIdx.x & 63, rg = (threadIdx.x >> 6) * 16;
    const int c = c0 + cc;
    float s = 0.f;
    __align__(16) __nv_bfloat16 col[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = tile[rg + i][cc];
      s += v;
      col[i] = __float2bfloat16(v);
    }
    psum[threadIdx.x >> 6][cc] = s;
    if (dstT != nullptr && c < C) {
      const int r = r0 + rg;
      __nv_bfloat16* d = dstT + (int64_t)c * tpitch + r;
      if (r + 16 <= tpitch && (tpitch & 7) == 0 && ((reinterpret_cast<uintptr_t>(dstT) & 15) == 0)) {
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(&col[0]);
        *reinterpret_cast<uint4*>(d + 8) = *reinterpret_cast<const uint4*>(&col[8]);
      } else {
        for (int i = 0; i < 16; ++i)
          if (r + i < tpitch) d[i] = col[i];
      }
    }
  }
  if (part != nullptr) {
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < C)
      part[(int64_t)blockIdx.x * C + c0 + threadIdx.x] =
          (psum[0][threadIdx.x] + psum[1][threadIdx.x]) + (psum[2][threadIdx.x] + psum[3][threadIdx.x]);
  }
}

extern "C" int64_t xdfm_cvt_bf16_both_workspace_bytes(int R, int C) { return (int64_t)((R + 63) / 64) * C * 4; }

extern "C" int xdfm_cvt_bf16_both(const float* src, const float* y, int act, int R, int C, int64_t ld, void* dst, int64_t dst_pitch,
                                  void* dstT, int64_t dstT_pitch, float* colsum, void* workspace, int64_t workspace_bytes, void* stream) {
  XDFM_CHECK_ARG(R >= 0 && C >= 1 && ld >= C, "cvt_bf16_both: bad shape R=%d C=%d ld=%lld", R, C, (long long)ld);
  XDFM_CHECK_ARG(dst == nullptr || (dst_pitch % 8 == 0 && dst_pitch >= C), "cvt_bf16_both: dst_pitch=%lld", (long long)dst_pitch);
  XDFM_CHECK_ARG(dstT == nullptr || (dstT_pitch % 8 == 0 && dstT_pitch >= R), "cvt_bf16_both: dstT_pitch=%lld", (long long)dstT_pitch);
  XDFM_CHECK_ARG(colsum == nullptr || (workspace != nullptr && workspace_bytes >= xdfm_cvt_bf16_both_workspace_bytes(R, C)),
                 "cvt_bf16_both: workspace too small for the column sums");
  if (R == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // the tile grid covers the padding of both outputs (zeros)
  const int64_t cover_r = std::max<int64_t>(R, dstT != nullptr ? dstT_pitch : 0);
  const int64_t cover_c = std::max<int64_t>(C, dst != nullptr ? dst_pitch : 0);
  dim3 grid((unsigned)ceil_div64(cover_r, 64), (unsigned)ceil_div64(cover_c, 64));
  cvt_bf16_both_kernel<<<grid, 256, 0, st>>>(src, y, act, R, C, ld, (__nv_bfloat16*)dst, dst_pitch, (__nv_bfloat16*)dstT, dstT_pitch,
                                             colsum != nullptr ? (float*)workspace : nullptr);
  XDFM_LAUNCH_CHECK();
  if (colsum != nullptr) {
    XDFM_TILE_COLSUM((const float*)workspace, (int64_t)((R + 63) / 64), (int64_t)C, C, colsum, 0, st);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}
