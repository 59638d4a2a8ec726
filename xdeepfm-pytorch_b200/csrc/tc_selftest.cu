// Minimal tcgen05 self-test: one CTA computes D[128, N] = A[128, K] * B[N, K]^T (bf16 in, fp32 out) with
//   mode 0: A from shared memory (TMA, SWIZZLE_128B, K-major)        -- tcgen05.mma  SS
//   mode 1: A written into TMEM by the threads (tcgen05.st, packed)  -- tcgen05.mma  TS
//   mode 2: as 1, K tail of B in SWIZZLE_32B boxes
//   mode 3: as 1, B given TRANSPOSED ([K, N] row-major, N contiguous) and consumed MN-major: [K rows x 64 columns] SWIZZLE_128B
//           boxes, descriptor LBO = box stride, SBO = 8 K-rows (the layout the dW kernels read row-major activations in)
// It exercises exactly the primitives the CIN kernels are built from (TMEM alloc, TMA swizzle vs. UMMA descriptor,
// A-in-TMEM layout, commit/mbarrier, tcgen05.ld) so that a descriptor/layout mistake shows up in a 40-line kernel.
#include "tc_common.cuh"
#include "../../include/xdfm.h"
#include <cudaTypedefs.h>

using namespace tc;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix (row pitch in bytes), box [box_rows, box_cols];
// swizzle: 0 = none, 1 = 128-byte (box_cols must then be 64), 2 = 32-byte (box_cols must then be 16)
int xdfm_make_tmap_bf16(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_rows,
                        uint32_t box_cols, int swizzle128) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  if (enc == nullptr) {
    xdfm_set_error("cuTensorMapEncodeTiled entry point not available");
    return XDFM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    xdfm_set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu box=%ux%u", (int)r, (unsigned long long)rows,
                   (unsigned long long)cols, (unsigned long long)row_pitch_bytes, box_rows, box_cols);
    return XDFM_ERR_CUDA;
  }
  return XDFM_OK;
}

int xdfm_make_tmap_bf16_sw128(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                              uint32_t box_rows) {
  return xdfm_make_tmap_bf16(out, gptr, rows, cols, row_pitch_bytes, box_rows, 64, 1);
}

__global__ void __launch_bounds__(128) tc_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ CUtensorMap tmBt, const __nv_bfloat16* __restrict__ A,
                                                          int N, int K, int mode, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = mode == 3 ? (N + 63) / 64 : K / 64;      // mode 3: chunks run along N
  const int ntail = mode == 3 ? 0 : (K % 64) / 16;                     // mode 2 only: 16-wide K steps past the last full chunk, SWIZZLE_32B boxes
  uint8_t* sB = smem;                                  // nchunk x [N rows x 128 B]   (mode 3: nchunk x [K rows x 128 B])
  uint8_t* sBt = smem + (size_t)nchunk * N * 128;      // ntail x [N rows x 32 B]
  uint8_t* sA = sBt + (size_t)ntail * N * 32;          // nchunk x [128 rows x 128 B]  (mode 0)
  sA = (uint8_t*)(((uintptr_t)sA + 1023) & ~(uintptr_t)1023);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t acc_col = 0, a_col = 256;
  if (tid == 0) {
    uint32_t bytes = (uint32_t)nchunk * N * 128 + (uint32_t)ntail * N * 32 + (mode == 0 ? (uint32_t)nchunk * 128 * 128 : 0u);
    if (mode == 3) bytes = (uint32_t)nchunk * K * 128;
    mbar_arrive_expect_tx(&bar_tma, bytes);
    for (int c = 0; c < nchunk && mode == 3; ++c) tma_load_2d(sB + (size_t)c * K * 128, &tmB, c * 64, 0, &bar_tma);
    for (int c = 0; c < nchunk && mode != 3; ++c) {
      tma_load_2d(sB + (size_t)c * N * 128, &tmB, c * 64, 0, &bar_tma);
      if (mode == 0) tma_load_2d(sA + (size_t)c * 128 * 128, &tmA, c * 64, 0, &bar_tma);
    }
    for (int t = 0; t < ntail; ++t) tma_load_2d(sBt + (size_t)t * N * 32, &tmBt, nchunk * 64 + t * 16, 0, &bar_tma);
  }
  if (mode >= 1) {
    // thread = row; pack (k, k+1) pairs into 32-bit columns
    const int row = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t r[4];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)row * K + k0);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = src[i];
      tmem_st_x4(tmem_base + lane_base + a_col + k0 / 2, r);
    }
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_wait(&bar_tma, 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, N, 0, mode == 3 ? 1 : 0);
    for (int ks = 0; ks < K / 16 && mode == 3; ++ks)
      umma_ts(tmem_base + acc_col, tmem_base + a_col + ks * 8, make_desc_mn_sw128(smem_u32(sB) + ks * 2048, (uint32_t)K * 128), idesc, ks > 0);
    for (int ks = 0; ks < K / 16 && mode != 3; ++ks) {
      int c = ks / 4, o = (ks % 4) * 32;
      uint64_t bdesc = c < nchunk ? make_desc_k_sw128(smem_u32(sB + (size_t)c * N * 128) + o)
                                  : make_desc_k_sw32(smem_u32(sBt + (size_t)(ks - nchunk * 4) * N * 32));
      if (mode == 0) {
        uint64_t adesc = make_desc_k_sw128(smem_u32(sA + (size_t)c * 128 * 128) + o);
        umma_ss(tmem_base + acc_col, adesc, bdesc, idesc, ks > 0);
      } else {
        umma_ts(tmem_base + acc_col, tmem_base + a_col + ks * 8, bdesc, idesc, ks > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  {
    const int row = tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + lane_base + acc_col + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) out[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
  (void)lane;
}

// A [128, K] bf16 row-major, Bm [N, K] bf16 row-major (both device), out [128, N] fp32; N % 16 == 0, N <= 256, K <= 256;
// modes 0 / 1: K % 64 == 0; mode 2 (A in TMEM, K tail of B in SWIZZLE_32B boxes): K % 16 == 0
extern "C" int xdfm_tc_selftest_gemm(const void* A, const void* Bm, int N, int K, int mode, float* out, void* stream) {
  XDFM_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K >= 16 && K <= 256 && (mode >= 2 ? K % 16 == 0 : (K % 64 == 0)),
                 "tc_selftest: bad N=%d K=%d mode=%d", N, K, mode);
  if (mode == 3) {
    // Bm is [K, N] row-major (row pitch N elements, N % 8 == 0): boxes of [K rows x 64 columns], columns past N read as zeros
    CUtensorMap tmT;
    int rc3 = xdfm_make_tmap_bf16(&tmT, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)N * 2, (uint32_t)K, 64, 1);
    if (rc3) return rc3;
    size_t sm3 = (size_t)((N + 63) / 64) * K * 128 + 2048;
    XDFM_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
    tc_selftest_kernel<<<1, 128, sm3, (cudaStream_t)stream>>>(tmT, tmT, tmT, (const __nv_bfloat16*)A, N, K, mode, out);
    XDFM_LAUNCH_CHECK();
    return XDFM_OK;
  }
  CUtensorMap tmA, tmB, tmBt;
  int rc = xdfm_make_tmap_bf16_sw128(&tmA, A, 128, K, (uint64_t)K * 2, 128);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16_sw128(&tmB, Bm, N, K, (uint64_t)K * 2, N);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmBt, Bm, N, K, (uint64_t)K * 2, N, 16, 2);
  if (rc) return rc;
  size_t sm = (size_t)(K / 64) * N * 128 + (size_t)((K % 64) / 16) * N * 32 + (size_t)(K / 64) * 128 * 128 + 2048;
  XDFM_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  tc_selftest_kernel<<<1, 128, sm, (cudaStream_t)stream>>>(tmA, tmB, tmBt, (const __nv_bfloat16*)A, N, K, mode, out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}


// ------------------------------------------------------------------------------------------------
// Latency probe (profiling only): how long do the hand-off primitives of the CIN kernels take on this chip?  One CTA, 320 threads:
// thread 0 issues, warps 2..9 are the "other side".  out[i] in SM cycles (clock64):
//   0  tcgen05.commit with nothing in flight -> own mbarrier wait returns
//   1  mbarrier.arrive (own thread) -> own wait returns
//   2  13 x tcgen05.mma (M128 N112 K16, TS) + commit -> wait returns
//   3  the same 13 MMAs, issue only (time the issuing thread is held)
//   4  try_wait on an already completed phase
//   5  tcgen05.ld 32x32b.x16 + wait::ld
//   6  round trip: thread 0 arrives on A, lane 0 of warp 2 waits on A and arrives on B, thread 0 waits on B
//   7  the same with tcgen05.commit on thread 0's side
//   8  the same with TWO commits per trip (the second barrier is the one waited on)
//   9  round trip by arrive against EIGHT full warps (all 256 lanes wait on A, lane 0 of each arrives on B, count 8)
//  10  the same by commit
//  11  the same by two commits, the whole of warp 0 (32 lanes) waiting on B like the MMA warp does
//  12  as 11 plus the eight warps run tcgen05.fence / wait::ld / __syncwarp / a shared-memory store around the arrive (the skeleton of
//      the dX row warps)
__global__ void __launch_bounds__(384, 1) tc_latency_probe_kernel(long long* out, int setmax) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4 + 3 * 7];
  __shared__ uint32_t tmem_base_s;
  __shared__ float scratch[384];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 112 * 128 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0u;       // one SWIZZLE_128B B chunk of zeros
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    for (int md = 0; md < 7; ++md) {
      mbar_init(&bar[4 + 3 * md], 1);
      mbar_init(&bar[5 + 3 * md], md >= 3 ? 8 : 1);
      mbar_init(&bar[6 + 3 * md], 1);
    }
    fence_barrier_init();
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int REPS = 16;
  if (setmax) {                                   // the register hand-over of the CIN kernels: does it change the hand-off latencies?
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  }
  if (tid == 0) {
    uint32_t ph = 0;
    long long t0, acc;
    acc = 0;
    for (int r = 0; r < REPS; ++r) { t0 = clock64(); umma_commit(&bar[0]); mbar_wait(&bar[0], ph); acc += clock64() - t0; ph ^= 1; }
    out[0] = acc / REPS;
    acc = 0;
    for (int r = 0; r < REPS; ++r) { t0 = clock64(); mbar_arrive(&bar[0]); mbar_wait(&bar[0], ph); acc += clock64() - t0; ph ^= 1; }
    out[1] = acc / REPS;
    const uint32_t idesc = make_idesc_bf16(128, 112);
    const uint64_t bdesc = make_desc_k_sw128(smem_u32(smem));
    long long acc_issue = 0;
    acc = 0;
    for (int r = 0; r < REPS; ++r) {
      t0 = clock64();
      for (int ks = 0; ks < 13; ++ks) umma_ts(tb + 256, tb + (ks & 3) * 8, bdesc + (uint64_t)((ks & 3) * 2), idesc, ks > 0);
      long long t1 = clock64();
      umma_commit(&bar[0]);
      mbar_wait(&bar[0], ph);
      acc += clock64() - t0;
      acc_issue += t1 - t0;
      ph ^= 1;
    }
    out[2] = acc / REPS;
    out[3] = acc_issue / REPS;
    acc = 0;
    for (int r = 0; r < REPS; ++r) { t0 = clock64(); mbar_wait(&bar[0], ph ^ 1); acc += clock64() - t0; }
    out[4] = acc / REPS;
  }
  __syncwarp();
  if (warp == 0) {
    long long t0 = clock64();
    uint32_t v[16];
    for (int r = 0; r < REPS; ++r) {
      tmem_ld_x16(tb + 256, v);
      tmem_wait_ld();
      asm volatile("" ::"r"(v[0]), "r"(v[15]));
    }
    if (tid == 0) out[5] = (clock64() - t0) / REPS;
  }
  __syncthreads();
  // round trips: thread 0 (or the whole of warp 0) against one lane / eight warps
  for (int mode = 0; mode < 7; ++mode) {
    const bool many = mode >= 3;                 // eight full warps answer
    const int ncommit = (mode == 1 || mode == 4) ? 1 : ((mode == 2 || mode >= 5) ? 2 : 0);
    const bool warp0_waits = mode >= 5;
    const bool dress = mode == 6;
    uint64_t* bA = &bar[4 + 3 * mode];
    uint64_t* bB = &bar[5 + 3 * mode];
    uint64_t* bC = &bar[6 + 3 * mode];
    __syncthreads();
    if (warp == 0) {
      long long t0 = clock64();
      for (int r = 0; r < REPS; ++r) {
        if (tid == 0) {
          if (ncommit == 0) mbar_arrive(bA);
          if (ncommit == 2) umma_commit(bC);
          if (ncommit >= 1) umma_commit(bA);
        }
        __syncwarp();
        if (warp0_waits || tid == 0) mbar_wait(bB, r & 1);
        if (warp0_waits) fence_after_sync();
        __syncwarp();
      }
      if (tid == 0) out[6 + mode] = (clock64() - t0) / REPS;
    } else if (warp >= 4 && (many || (warp == 4 && lane == 0))) {
      for (int r = 0; r < REPS; ++r) {
        mbar_wait(bA, r & 1);
        if (dress) {
          fence_after_sync();
          tmem_wait_ld();
          scratch[tid] = (float)r;
          fence_before_sync();
        }
        if (many) __syncwarp();
        if (lane == 0) mbar_arrive(bB);
      }
    }
  }
  __syncthreads();
  // 13 / 14: the 13 MMAs + commit -> wait of test 2 while the eight other warps keep reading ANOTHER accumulator region out of TMEM
  // (13: tcgen05.ld x16 x 7 + wait::ld in a loop, the dX drain's traffic; 14: the same loop with the loads replaced by FMAs)
  __shared__ volatile int stop_flag;
  for (int mode = 0; mode < 2; ++mode) {
    if (tid == 0) stop_flag = 0;
    __syncthreads();
    if (tid == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 112);
      const uint64_t bdesc = make_desc_k_sw128(smem_u32(smem));
      uint32_t ph = 0;
      long long acc = 0;
      for (int r = 0; r < REPS; ++r) {
        long long t0 = clock64();
        for (int ks = 0; ks < 13; ++ks) umma_ts(tb + 256, tb + (ks & 3) * 8, bdesc + (uint64_t)((ks & 3) * 2), idesc, ks > 0);
        umma_commit(&bar[3]);
        mbar_wait(&bar[3], ph);
        acc += clock64() - t0;
        ph ^= 1;
      }
      out[13 + mode] = acc / REPS;
      stop_flag = 1;
    } else if (warp >= 4) {
      const uint32_t taddr = tb + 384 + ((uint32_t)((warp & 3) * 32) << 16);
      float accf = 0.f;
      while (!stop_flag) {
        uint32_t v[16];
        if (mode == 0) {
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            tmem_ld_x16(taddr + c * 16, v);
            tmem_wait_ld();
            accf += __uint_as_float(v[0]) + __uint_as_float(v[15]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 112; ++c) accf = fmaf(accf, 1.0001f, 0.5f);
        }
      }
      scratch[tid] = accf;
    }
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

extern "C" int xdfm_tc_latency_probe(long long* out, int setmaxnreg, void* stream) {
  tc_latency_probe_kernel<<<1, 384, 112 * 128, (cudaStream_t)stream>>>(out, setmaxnreg);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
