// CIN layer on the 5th-gen tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces (reference, file:line): deepctr/layers/interaction.py:218-246 -- einsum outer product -> Conv1d(k=1) -> ReLU ->
// split-half -> sum over D.  The reference materialises Z = X^{k-1} (x) X^0 as [B, h*m, D] fp32 in HBM and runs a batched GEMM
// with N = D.  Here the contraction is ONE implicit GEMM per layer
//
//        Y[n, h] = sum_k Z[n, k] * W'[h, k],      n = (sample, d) -> M rows,   h -> N columns,   k = (j, i) -> K
//
// whose A operand Z never exists in memory: 128 producer threads (one per row n) hold X^{k-1}[n, :] in registers, multiply by
// X^0[n, j] and write packed bf16 pairs straight into TENSOR MEMORY with tcgen05.st; tcgen05.mma (kind::f16, M=128,
// N=H_pad<=256, K=16) reads A from TMEM and B = W' (bf16, K-major, SWIZZLE_128B) from shared memory where TMA put it; the fp32
// accumulator [128 x H_pad] lives in TMEM and the epilogue (tcgen05.ld -> bias -> ReLU -> split/pool/store) reads it back.
//
// K ordering is private to this file: k' = j * HpP + i (j over X^0 fields, i over X^{k-1} channels padded to a multiple of 8),
// W' = cin_prep_w(W) is the reference weight permuted/padded/converted accordingly each step.
//
// TMEM columns (512 allocated): [0, 256) accumulator, [256, 384) and [384, 512) two A stages of up to 256 K-values each.
// Warp roles (10 warps): 0 = TMA (W' chunks, x tiles), 1 = MMA issuer + TMEM alloc, 2..5 = Z producers, 6..9 = epilogue.
#include "tc_common.cuh"
#include "../../include/xdfm.h"

using namespace tc;

#define TC_THREADS 320
#define TC_A_COL0 256
#define TC_A_STAGE_COLS 128
#define TC_NS_W 4          // W' chunk ring depth

struct CinTcParams {
  const __nv_bfloat16* x0b;   // [B, m, D]
  const __nv_bfloat16* xkb;   // element (b, i, d) at xkb[b*xk_bstride + i*D + d]
  int64_t xk_bstride;
  const float* bias;          // [H]
  __nv_bfloat16* yb;          // [B, H, D] post-activation (next layer's input / backward)
  float* pooled;              // [B, fm_total] or null
  float* maps;                // [B, fm_total, D] or null
  int64_t B;
  int m, Hp, H, H_pad, D, act, hdb, fm_total, col_off;
  int TB;                     // samples per tile = 128 / D
  int64_t n_tiles;
  int G;                      // X^0 fields per A stage
  int n_stages;               // A stages per tile
  int ksteps_total;           // UMMA K-steps per tile
  int n_wchunks;              // 64-wide W' chunks per tile
};

struct __align__(8) CinTcBars {
  uint64_t w_full[TC_NS_W], w_empty[TC_NS_W];
  uint64_t a_full[2], a_empty[2];
  uint64_t x_full[2], x_empty[2];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  __nv_bfloat162 v = __halves2bfloat162(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int NREG>
__device__ __forceinline__ void tmem_st_regs(uint32_t taddr, const uint32_t* r) {
  // NREG is a multiple of 4; emit the widest power-of-two stores
  if constexpr (NREG >= 16) {
    tmem_st_x16(taddr, r);
    tmem_st_regs<NREG - 16>(taddr + 16, r + 16);
  } else if constexpr (NREG >= 8) {
    tmem_st_x8(taddr, r);
    tmem_st_regs<NREG - 8>(taddr + 8, r + 8);
  } else if constexpr (NREG >= 4) {
    tmem_st_x4(taddr, r);
    tmem_st_regs<NREG - 4>(taddr + 4, r + 4);
  }
}

struct EpiRow {
  __nv_bfloat16* yb;   // &yb[b, 0, d] or null
  float* maps;         // &maps[b, col_off - hdb, d] or null (index by h)
  float* pooled;       // &pooled[b, col_off - hdb] or null (index by h)
  float* spool;        // per-warp scratch [H_pad] (D > 32)
  bool valid;
};

// One 16-column accumulator chunk of one row: bias + activation, bf16 / fp32 stores, and the sum over the sample's D lanes.
// The lane-sum uses a transposing butterfly: every shuffle level halves the number of live values per lane, so 16 columns
// cost 8+4+2+1 shuffles instead of 16*log2(LD); afterwards lane l holds the finished sum of column `col_of_lane`.
template <int LD>   // lanes per sample inside the warp: min(D, 32)
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[16], int c0, int lane, const float* __restrict__ sBias,
                                               const CinTcParams& p, const EpiRow& r) {
  float y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float t = __uint_as_float(v[i]) + sBias[c0 + i];
    if (p.act == XDFM_ACT_RELU) t = fmaxf(t, 0.f);
    else if (p.act == XDFM_ACT_SIGMOID) t = 1.f / (1.f + __expf(-t));
    y[i] = r.valid ? t : 0.f;
  }
  const int hmax = p.H - c0;   // columns >= hmax are padding
  if (r.yb != nullptr) {
    __nv_bfloat16* yp = r.yb + (int64_t)c0 * p.D;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < hmax) yp[i * p.D] = __float2bfloat16(y[i]);
  }
  if (r.maps != nullptr) {
    float* mp = r.maps + (int64_t)c0 * p.D;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < hmax && c0 + i >= p.hdb) mp[(int64_t)i * p.D] = y[i];
  }
  if (p.pooled == nullptr) return;
  int nv = 16, col = 0;
#pragma unroll
  for (int o = LD / 2; o >= 1; o >>= 1) {
    if (nv > 1) {
      const bool upper = (lane & o) != 0;
      const int half = nv / 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < half) {
          const float send = upper ? y[i] : y[i + half];
          const float keep = upper ? y[i + half] : y[i];
          y[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      if (upper) col += half;
      nv = half;
    } else {
      y[0] += __shfl_xor_sync(0xffffffffu, y[0], o);
    }
  }
  // lane now owns columns c0 + col .. c0 + col + nv - 1 (LD=8: two columns, LD=16: one, LD=32: one shared by a lane pair)
  if (LD == 32 && (lane & 1)) return;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (i < nv) {
      const int h = c0 + col + i;
      if (p.D > 32) r.spool[h] = y[i];
      else if (r.pooled != nullptr && h >= p.hdb && h < p.H) r.pooled[h] = y[i];
    }
  }
}

// NI8 = HpP / 8 (channels of X^{k-1} padded to a multiple of 8, HpP <= 128)
template <int NI8>
__global__ void __launch_bounds__(TC_THREADS, 1) cin_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW, CinTcParams p) {
  constexpr int HpP = NI8 * 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- shared memory carve-up
  const uint32_t w_stage_bytes = (uint32_t)p.H_pad * 128;
  uint8_t* sW = smem;                                                       // TC_NS_W x [H_pad x 128 B], 1024-aligned
  const uint32_t x0_bytes = (uint32_t)p.TB * p.m * p.D * 2;
  const uint32_t xk_row_bytes = (uint32_t)p.Hp * p.D * 2;                   // one sample
  const uint32_t xk_bytes = (uint32_t)p.TB * xk_row_bytes;
  uint8_t* sX0 = sW + (size_t)TC_NS_W * w_stage_bytes;                      // 2 x [TB][m][D] bf16
  uint8_t* sXk = sX0 + 2 * (size_t)((x0_bytes + 127) & ~127u);              // 2 x [TB][Hp][D] bf16
  uint8_t* sEnd = sXk + 2 * (size_t)((xk_bytes + 127) & ~127u);
  float* sBias = reinterpret_cast<float*>(sEnd);                            // [H_pad]
  float* sPool = sBias + p.H_pad;                                           // [4][H_pad] cross-warp pooling scratch (D > 32)
  CinTcBars* bars = reinterpret_cast<CinTcBars*>(sPool + 4 * p.H_pad);
  const uint32_t x0_stride = (x0_bytes + 127) & ~127u, xk_stride = (xk_bytes + 127) & ~127u;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_NS_W; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 128); mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->x_full[i], 1);   mbar_init(&bars->x_empty[i], 128);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, 128);
    fence_barrier_init();
  }
  for (int h = threadIdx.x; h < p.H_pad; h += TC_THREADS) sBias[h] = h < p.H ? p.bias[h] : 0.f;
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  const int64_t tile0 = blockIdx.x, tstep = gridDim.x;

  if (warp == 0) {
    // =============================== TMA producer: x tiles (one tile ahead) and the W' chunk stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      auto load_x = [&](int64_t tile, int it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(&bars->x_empty[buf], ((it >> 1) - 1) & 1);
        const int64_t b0 = tile * p.TB;
        const int nb = (int)min((int64_t)p.TB, p.B - b0);
        mbar_arrive_expect_tx(&bars->x_full[buf], (uint32_t)nb * (uint32_t)(p.m * p.D * 2) + (uint32_t)nb * xk_row_bytes);
        bulk_load_1d(sX0 + (size_t)buf * x0_stride, p.x0b + b0 * p.m * p.D, (uint32_t)nb * (uint32_t)(p.m * p.D * 2), &bars->x_full[buf]);
        for (int s = 0; s < nb; ++s)
          bulk_load_1d(sXk + (size_t)buf * xk_stride + (size_t)s * xk_row_bytes, p.xkb + (b0 + s) * p.xk_bstride, xk_row_bytes,
                       &bars->x_full[buf]);
      };
      int it = 0;
      uint32_t wc = 0;  // global W' chunk counter (ring position)
      if (tile0 < p.n_tiles) load_x(tile0, 0);
      for (int64_t tile = tile0; tile < p.n_tiles; tile += tstep, ++it) {
        if (tile + tstep < p.n_tiles) load_x(tile + tstep, it + 1);
        for (int c = 0; c < p.n_wchunks; ++c, ++wc) {
          const int ws = wc % TC_NS_W;
          if (wc >= TC_NS_W) mbar_wait(&bars->w_empty[ws], ((wc / TC_NS_W) - 1) & 1);
          mbar_arrive_expect_tx(&bars->w_full[ws], w_stage_bytes);
          tma_load_2d(sW + (size_t)ws * w_stage_bytes, &tmW, c * 64, 0, &bars->w_full[ws]);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.H_pad);
      uint32_t wc = 0, sc = 0;  // global W' chunk / A stage counters
      int it = 0;
      for (int64_t tile = tile0; tile < p.n_tiles; tile += tstep, ++it) {
        if (it > 0) {
          mbar_wait(&bars->acc_empty, (it - 1) & 1);
          fence_after_sync();
        }
        int ks = 0;  // K-step within the tile
        for (int s = 0; s < p.n_stages; ++s, ++sc) {
          const int sb = sc & 1;
          mbar_wait(&bars->a_full[sb], (sc >> 1) & 1);
          fence_after_sync();
          const int nj = min(p.G, p.m - s * p.G);
          const int nks = (nj * HpP + 15) / 16;
          const uint32_t a_col = TC_A_COL0 + sb * TC_A_STAGE_COLS;
          for (int t = 0; t < nks; ++t, ++ks) {
            const int ws = wc % TC_NS_W;
            if ((ks & 3) == 0) {
              mbar_wait(&bars->w_full[ws], (wc / TC_NS_W) & 1);
              fence_after_sync();
            }
            const uint64_t bdesc = make_desc_k_sw128(smem_u32(sW + (size_t)ws * w_stage_bytes) + (ks & 3) * 32);
            umma_ts(tmem_base, tmem_base + a_col + t * 8, bdesc, idesc, ks > 0);
            if ((ks & 3) == 3 || ks == p.ksteps_total - 1) {
              umma_commit(&bars->w_empty[ws]);   // W' chunk consumed
              ++wc;
            }
          }
          umma_commit(&bars->a_empty[sb]);       // A stage consumed
        }
        umma_commit(&bars->acc_full);            // accumulator complete
      }
    }
  } else if (warp < 6) {
    // =============================== Z producers: thread <-> accumulator row n = (sample bl, d) ===============================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int n = q * 32 + lane;
    const int bl = n / p.D, d = n - bl * p.D;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t sc = 0;
    int it = 0;
    for (int64_t tile = tile0; tile < p.n_tiles; tile += tstep, ++it) {
      const int buf = it & 1;
      mbar_wait(&bars->x_full[buf], (it >> 1) & 1);
      const __nv_bfloat16* x0s = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)buf * x0_stride) + (size_t)bl * p.m * p.D + d;
      const __nv_bfloat16* xks = reinterpret_cast<const __nv_bfloat16*>(sXk + (size_t)buf * xk_stride) + (size_t)bl * p.Hp * p.D + d;
      // X^{k-1}[n, 0..HpP) as packed bf16 pairs (zero padded)
      __nv_bfloat162 xk2[HpP / 2];
#pragma unroll
      for (int i2 = 0; i2 < HpP / 2; ++i2) {
        const int i = 2 * i2;
        __nv_bfloat16 lo = (i < p.Hp && bl < p.TB) ? xks[(size_t)i * p.D] : __float2bfloat16(0.f);
        __nv_bfloat16 hi = (i + 1 < p.Hp && bl < p.TB) ? xks[(size_t)(i + 1) * p.D] : __float2bfloat16(0.f);
        xk2[i2] = __halves2bfloat162(lo, hi);
      }
      for (int s = 0; s < p.n_stages; ++s, ++sc) {
        const int sb = sc & 1;
        if (sc >= 2) {
          mbar_wait(&bars->a_empty[sb], ((sc >> 1) - 1) & 1);
          fence_after_sync();
        }
        const int nj = min(p.G, p.m - s * p.G);
        uint32_t col = tmem_base + lane_addr + TC_A_COL0 + sb * TC_A_STAGE_COLS;
        for (int jj = 0; jj < nj; ++jj) {
          const int j = s * p.G + jj;
          const __nv_bfloat16 xv = (bl < p.TB) ? x0s[(size_t)j * p.D] : __float2bfloat16(0.f);
          const __nv_bfloat162 xv2 = __halves2bfloat162(xv, xv);
          uint32_t z[HpP / 2];
#pragma unroll
          for (int i2 = 0; i2 < HpP / 2; ++i2) {
            __nv_bfloat162 prod = __hmul2(xk2[i2], xv2);
            z[i2] = *reinterpret_cast<uint32_t*>(&prod);
          }
          tmem_st_regs<HpP / 2>(col, z);
          col += HpP / 2;
        }
        if ((nj * HpP) & 15) {   // odd tail: pad the last K-step of the stage with zeros
          uint32_t zz[4] = {0u, 0u, 0u, 0u};
          tmem_st_x4(col, zz);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars->a_full[sb]);
      }
      mbar_arrive(&bars->x_empty[buf]);
    }
  } else {
    // =============================== epilogue: TMEM -> bias -> activation -> y / pooled / maps ===============================
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const int bl = n / p.D, d = n - bl * p.D;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int ew = warp - 6;                     // 0..3 (epilogue warp index, for the cross-warp pooling scratch)
    int it = 0;
    for (int64_t tile = tile0; tile < p.n_tiles; tile += tstep, ++it) {
      const int64_t b = tile * p.TB + bl;
      const bool valid = bl < p.TB && b < p.B;
      EpiRow r;
      r.valid = valid;
      r.yb = (p.yb && valid) ? p.yb + b * (int64_t)p.H * p.D + d : nullptr;
      r.maps = (p.maps && valid) ? p.maps + (b * p.fm_total + p.col_off - p.hdb) * (int64_t)p.D + d : nullptr;
      r.pooled = (p.pooled && valid) ? p.pooled + b * p.fm_total + p.col_off - p.hdb : nullptr;
      r.spool = sPool + ew * p.H_pad;
      mbar_wait(&bars->acc_full, it & 1);
      fence_after_sync();
      for (int c0 = 0; c0 < p.H_pad; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + lane_addr + c0, v);
        tmem_wait_ld();
        switch (p.D) {
          case 8: epilogue_chunk<8>(v, c0, lane, sBias, p, r); break;
          case 16: epilogue_chunk<16>(v, c0, lane, sBias, p, r); break;
          default: epilogue_chunk<32>(v, c0, lane, sBias, p, r); break;   // D = 32, 64, 128: full-warp segments
        }
      }
      fence_before_sync();
      mbar_arrive(&bars->acc_empty);             // accumulator drained: the next tile's MMAs may start
      if (p.pooled && p.D > 32) {
        // D = 64: warps (0,1) and (2,3) hold one sample each; D = 128: all four warps hold one sample
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int wps = p.D / 32;                // warps per sample
        for (int e = threadIdx.x - 6 * 32; e < (4 / wps) * p.H_pad; e += 128) {
          const int sidx = e / p.H_pad, h = e - sidx * p.H_pad;
          const int64_t bb = tile * p.TB + sidx;
          if (bb < p.B && h >= p.hdb && h < p.H) {
            float s = 0.f;
            for (int w2 = 0; w2 < wps; ++w2) s += sPool[(((sidx * wps + w2) + 2) & 3) * p.H_pad + h];
            p.pooled[bb * p.fm_total + p.col_off + (h - p.hdb)] = s;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// W fp32 [H, Hp*m] (k = i*m + j) -> W' bf16 [H_pad, KP] (k' = j*HpP + i), zero padded
// ------------------------------------------------------------------------------------------------
__global__ void cin_prep_w_kernel(const float* __restrict__ W, int H, int Hp, int m, int HpP, int H_pad, int KP,
                                  __nv_bfloat16* __restrict__ Wp) {
  int64_t total = (int64_t)H_pad * KP;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int h = (int)(e / KP), kp = (int)(e - (int64_t)h * KP);
    int j = kp / HpP, i = kp - j * HpP;
    float v = 0.f;
    if (h < H && j < m && i < Hp) v = W[(int64_t)h * Hp * m + (int64_t)i * m + j];
    Wp[e] = __float2bfloat16(v);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

extern "C" int xdfm_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  f32_to_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct CinTcGeom {
  int HpP, H_pad, KP, G, n_stages, ksteps_total, n_wchunks, TB;
  size_t smem;
};

static int cin_tc_geom(int m, int Hp, int H, int D, CinTcGeom* g) {
  if (!(D == 8 || D == 16 || D == 32 || D == 64 || D == 128)) {
    xdfm_set_error("cin_tc: embedding dim D=%d unsupported by the bf16 tensor-core path (needs 8/16/32/64/128); use cin_precision='fp32'", D);
    return XDFM_ERR_UNSUPPORTED;
  }
  if (Hp > 128 || H > 256 || m > XDFM_MAX_FIELDS) {
    xdfm_set_error("cin_tc: layer too wide for the bf16 tensor-core path (Hp=%d<=128, H=%d<=256, m=%d<=64); use cin_precision='fp32'", Hp, H, m);
    return XDFM_ERR_UNSUPPORTED;
  }
  g->HpP = round_up(Hp, 8);
  g->H_pad = round_up(H, 16);
  g->G = 256 / g->HpP;
  if ((g->HpP % 16) != 0) g->G -= (g->G & 1);
  if (g->G < 1) g->G = 1;
  g->n_stages = (m + g->G - 1) / g->G;
  int ks = 0;
  for (int s = 0; s < g->n_stages; ++s) {
    int nj = std::min(g->G, m - s * g->G);
    ks += (nj * g->HpP + 15) / 16;
  }
  g->ksteps_total = ks;
  g->n_wchunks = (ks + 3) / 4;
  g->KP = g->n_wchunks * 64;
  g->TB = 128 / D;
  size_t x0 = ((size_t)g->TB * m * D * 2 + 127) & ~(size_t)127;
  size_t xk = ((size_t)g->TB * Hp * D * 2 + 127) & ~(size_t)127;
  g->smem = (size_t)TC_NS_W * g->H_pad * 128 + 2 * x0 + 2 * xk + (size_t)5 * g->H_pad * 4 + sizeof(CinTcBars) + 1024;
  if (g->smem > 227 * 1024) {
    xdfm_set_error("cin_tc: shared memory %zu exceeds 227 KB (m=%d Hp=%d H=%d D=%d)", g->smem, m, Hp, H, D);
    return XDFM_ERR_UNSUPPORTED;
  }
  return XDFM_OK;
}

extern "C" int64_t xdfm_cin_tc_wprime_elems(int m, int Hp, int H, int D) {
  CinTcGeom g;
  if (cin_tc_geom(m, Hp, H, D, &g) != XDFM_OK) return -1;
  return (int64_t)g.H_pad * g.KP;
}

template <int NI8>
static int launch_cin_fwd_tc(const CUtensorMap& tm, const CinTcParams& p, size_t smem, int blocks, cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_fwd_tc_kernel<NI8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cin_fwd_tc_kernel<NI8><<<blocks, TC_THREADS, smem, st>>>(tm, p);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// x0b [B,m,D] bf16; xkb bf16 (b,i,d) at xkb[b*xk_bstride + i*D + d]; W fp32 [H, Hp*m]; wprime = scratch bf16 [xdfm_cin_tc_wprime_elems]
// yb [B,H,D] bf16 out; pooled / maps as in xdfm_cin_fwd_f32.
extern "C" int xdfm_cin_fwd_tc(const void* x0b, const void* xkb, int64_t xk_bstride, const float* W, const float* bias, void* wprime,
                               int64_t B, int m, int Hp, int H, int D, int act, void* yb, int direct_begin, float* pooled, float* maps,
                               int fm_total, int col_off, void* stream) {
  CinTcGeom g;
  int rc = cin_tc_geom(m, Hp, H, D, &g);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CHECK_ARG(((uintptr_t)x0b % 16 == 0) && ((uintptr_t)xkb % 16 == 0) && (xk_bstride * 2) % 16 == 0 && (Hp * D * 2) % 16 == 0,
                 "cin_fwd_tc: x tiles must be 16-byte aligned (Hp*D*2=%d)", Hp * D * 2);
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t total = (int64_t)g.H_pad * g.KP;
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64(total, 256));
    // the K order inside a stage is contiguous in k' = j*HpP + i, so the padded column index equals k'
    cin_prep_w_kernel<<<blocks, 256, 0, st>>>(W, H, Hp, m, g.HpP, g.H_pad, g.KP, (__nv_bfloat16*)wprime);
    XDFM_LAUNCH_CHECK();
  }
  CUtensorMap tm;
  rc = xdfm_make_tmap_bf16_sw128(&tm, wprime, (uint64_t)g.H_pad, (uint64_t)g.KP, (uint64_t)g.KP * 2, (uint32_t)g.H_pad);
  if (rc) return rc;
  CinTcParams p;
  p.x0b = (const __nv_bfloat16*)x0b; p.xkb = (const __nv_bfloat16*)xkb; p.xk_bstride = xk_bstride; p.bias = bias;
  p.yb = (__nv_bfloat16*)yb; p.pooled = pooled; p.maps = maps; p.B = B;
  p.m = m; p.Hp = Hp; p.H = H; p.H_pad = g.H_pad; p.D = D; p.act = act; p.hdb = direct_begin; p.fm_total = fm_total; p.col_off = col_off;
  p.TB = g.TB; p.n_tiles = ceil_div64(B, g.TB); p.G = g.G; p.n_stages = g.n_stages; p.ksteps_total = g.ksteps_total; p.n_wchunks = g.n_wchunks;
  int blocks = (int)std::min<int64_t>(p.n_tiles, xdfm_num_sms());
  switch (g.HpP / 8) {
#define CASE_NI8(n) case n: return launch_cin_fwd_tc<n>(tm, p, g.smem, blocks, st);
    CASE_NI8(1) CASE_NI8(2) CASE_NI8(3) CASE_NI8(4) CASE_NI8(5) CASE_NI8(6) CASE_NI8(7) CASE_NI8(8)
    CASE_NI8(9) CASE_NI8(10) CASE_NI8(11) CASE_NI8(12) CASE_NI8(13) CASE_NI8(14) CASE_NI8(15) CASE_NI8(16)
#undef CASE_NI8
  }
  xdfm_set_error("cin_fwd_tc: unreachable HpP=%d", g.HpP);
  return XDFM_ERR_UNSUPPORTED;
}
