// Shared helpers for libxdfm_sm100a.so (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>

#define XDFM_OK 0
#define XDFM_ERR_ARG 1
#define XDFM_ERR_CUDA 2
#define XDFM_ERR_UNSUPPORTED 3

void xdfm_set_error(const char* fmt, ...);

#define XDFM_CHECK_ARG(cond, ...)              \
  do {                                         \
    if (!(cond)) {                             \
      xdfm_set_error(__VA_ARGS__);             \
      return XDFM_ERR_ARG;                     \
    }                                          \
  } while (0)

#define XDFM_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      xdfm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return XDFM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

extern long long g_xdfm_launches;
#define XDFM_LAUNCH_CHECK()                                                                 \
  do {                                                                                      \
    ++g_xdfm_launches;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      xdfm_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return XDFM_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

static inline int xdfm_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    const char* dbg = getenv("XDFM_DEBUG_SMS");       // profiling experiments only: size persistent grids for fewer SMs
    if (dbg != nullptr && atoi(dbg) > 0 && atoi(dbg) < n) n = atoi(dbg);
  }
  return n;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[c] (+)= sum over the n_tiles rows of part[t * pitch + c]: second stage of every deterministic column sum in the library.
// 8 columns x 128 tile strides per block of 1024 threads; a thread has up to 16 loads in flight per pass and adds them in a fixed
// order, the 128 strides are then added as a tree in shared memory -- the result depends only on (n_tiles, values).
static __global__ void __launch_bounds__(1024) xdfm_tile_colsum_kernel(const float* __restrict__ part, int64_t n_tiles, int64_t pitch, int C,
                                                                        float* __restrict__ out, int accumulate) {
  __shared__ float red[128][9];
  const int cx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cx;
  float acc = 0.f;
  if (c < C) {
    for (int64_t t0 = ty; t0 < n_tiles; t0 += 128 * 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int64_t t = t0 + (int64_t)i * 128;
        v[i] = t < n_tiles ? part[t * pitch + c] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) acc += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
    }
  }
  red[ty][cx] = acc;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (ty < s) red[ty][cx] += red[ty + s][cx];
    __syncthreads();
  }
  if (ty == 0 && c < C) out[c] = accumulate ? out[c] + red[0][cx] : red[0][cx];
}
#define XDFM_TILE_COLSUM(part, n_tiles, pitch, C, out, accumulate, st) \
  xdfm_tile_colsum_kernel<<<(unsigned)(((C) + 7) / 8), 1024, 0, (st)>>>((part), (n_tiles), (pitch), (C), (out), (accumulate))

// 128-bit streaming loads/stores (read-only path, do not pollute L1)
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
