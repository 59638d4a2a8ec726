// CIN layer weight gradient on the tensor cores.
//
// Replaces the dW part of torch's autograd of deepctr/layers/interaction.py:224 (convolution_backward weight grad):
//
//     dW[h, i*m + j] = sum_r dY[r,h] * Xk[r,i] * X0[r,j]          r = (sample, d) over all B*D rows
//
// The reduction runs over r, so this kernel reads CHANNEL-MAJOR copies (xdfm_rows_to_cols_bf16): dyT [H_pad, R], xkT [HpQ, R],
// x0T [mP, R] (r contiguous).  For a field j:  dW_j^T[i, h] = sum_r (xkT[i,r] * x0T[j,r]) * dyT[h,r]  is a GEMM with
// M = i (128 TMEM lanes), N = h (H_pad), K = r, whose A operand (the scaled xk row) is produced on the fly: thread = lane i
// reads its xkT row chunk from (swizzled) shared memory, multiplies by the broadcast x0T[j] chunk and stores packed bf16 pairs
// into TMEM; B = dyT tile [H_pad x 64 r] comes from TMA (K-major, SWIZZLE_128B).  A CTA owns JP fields (JP accumulators of
// H_pad columns in TMEM) and one range of r; CTAs of a cluster share the r range (different fields), so the dyT / xkT tiles are
// fetched once per cluster and multicast.  Each CTA writes its fp32 partial [JP, HpQ, H_pad]; a second kernel sums the r-splits
// in fixed order and un-permutes into the reference layout [H, Hp*m] (deterministic).
//
// Narrow X^{k-1} (layer 0: Hp = m = 26 at Criteo shape): with one field per accumulator only Hp of the 128 TMEM lanes carry data
// and the tensor core spends 128 / Hp of the necessary MMA work.  When HpQ <= 64 the lanes are PACKED: lane = (jsub, i) with
// LW = 128 / PACK lanes per field (PACK = 4 for HpQ <= 32, 2 for HpQ <= 64), so ONE accumulator holds dW_j^T for PACK consecutive
// fields and a CTA owns JP * PACK fields.  A warp covers 32 lanes = one field (PACK = 4) or half a field (PACK = 2), so the x0T
// row a thread multiplies by is still a warp-wide broadcast.
#include "tc_common.cuh"
#include "../../include/xdfm.h"

using namespace tc;

#define DW_THREADS 192      // warps: 0 TMA, 1 MMA + TMEM alloc, 2..5 producers / epilogue
#define DW_MAX_NS 4
#define DW_MAX_ASLOTS 4

struct CinDwParams {
  float* part;                // [n_splits, m, HpQ, H_pad] fp32
  int64_t R;
  int m, Hp, HpQ, H_pad, JP;
  int PACK, LW;               // fields per accumulator (lane packing), lanes per field = 128 / PACK
  int n_jgroups;              // real field groups = ceil(m / (JP * PACK))
  int n_jgroups_padded;       // multiple of the cluster size
  int n_splits;
  int64_t chunks_total;       // ceil(R / 64)
  int64_t chunks_per_split;
  int ns;                     // stage ring depth
  int a_slots;                // A ring slots (32 TMEM columns each)
  int a_col0;                 // first TMEM column of the A ring
};

struct __align__(8) CinDwBars {
  uint64_t x_full[DW_MAX_NS], x_empty[DW_MAX_NS];
  uint64_t a_full[DW_MAX_ASLOTS], a_empty[DW_MAX_ASLOTS];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(DW_THREADS, 1)
cin_bwd_dw_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmXk,
                     const __grid_constant__ CUtensorMap tmX0, CinDwParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int FPC = p.JP * p.PACK;                                                   // fields per CTA
  const uint32_t dy_bytes = (uint32_t)p.H_pad * 128, xk_bytes = (uint32_t)p.HpQ * 128, x0_bytes = (uint32_t)FPC * 128;
  const uint32_t stage_bytes = dy_bytes + xk_bytes + 1024;                      // x0 rows live in the last 1 KB (keeps 1024-alignment)
  uint8_t* sStage = smem;
  CinDwBars* bars = reinterpret_cast<CinDwBars*>(smem + (size_t)p.ns * stage_bytes);

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);
  const int jgroup = blockIdx.x % p.n_jgroups_padded;
  const int split = blockIdx.x / p.n_jgroups_padded;
  const bool active = jgroup < p.n_jgroups;
  const int j0 = jgroup * FPC;
  const int nj = active ? min(p.JP, (p.m - j0 + p.PACK - 1) / p.PACK) : 0;        // accumulators in use (PACK fields each)
  const int64_t c_beg = (int64_t)split * p.chunks_per_split;
  const int64_t c_end = min(p.chunks_total, c_beg + p.chunks_per_split);
  const int n_chunks = (int)max((int64_t)0, c_end - c_beg);

  if (threadIdx.x == 0) {
    for (int i = 0; i < DW_MAX_NS; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], csize); }
    for (int i = 0; i < DW_MAX_ASLOTS; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], 1); }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // =============================== TMA ===============================
    if (lane == 0) {
      prefetch_tmap(&tmDy);
      prefetch_tmap(&tmXk);
      prefetch_tmap(&tmX0);
      const int dy_slice = p.H_pad / (int)csize, xk_slice = p.HpQ / (int)csize;
      uint32_t st = 0, phase = 1;
      bool first_pass = true;
      for (int c = 0; c < n_chunks; ++c) {
        if (!first_pass) mbar_wait(&bars->x_empty[st], phase);
        uint8_t* base = sStage + (size_t)st * stage_bytes;
        mbar_arrive_expect_tx(&bars->x_full[st], dy_bytes + xk_bytes + (active ? x0_bytes : 0u));
        const int rcol = (int)((c_beg + c) * 64);
        if (csize > 1) {
          tma_load_2d_mcast(base + (size_t)crank * dy_slice * 128, &tmDy, rcol, (int)crank * dy_slice, &bars->x_full[st], cmask);
          tma_load_2d_mcast(base + dy_bytes + (size_t)crank * xk_slice * 128, &tmXk, rcol, (int)crank * xk_slice, &bars->x_full[st], cmask);
        } else {
          tma_load_2d(base, &tmDy, rcol, 0, &bars->x_full[st]);
          tma_load_2d(base + dy_bytes, &tmXk, rcol, 0, &bars->x_full[st]);
        }
        if (active) tma_load_2d(base + dy_bytes + xk_bytes, &tmX0, rcol, j0, &bars->x_full[st]);   // this CTA's own field rows
        if (++st == (uint32_t)p.ns) { st = 0; phase ^= 1; first_pass = false; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = make_idesc_bf16(128, p.H_pad);
    uint32_t st = 0, phase = 0, as = 0, aphase = 0;
    for (int c = 0; c < n_chunks; ++c) {
      mbar_wait(&bars->x_full[st], phase);
      fence_after_sync();
      const uint64_t bdesc = make_desc_k_sw128(smem_u32(sStage + (size_t)st * stage_bytes));
      for (int jj = 0; jj < nj; ++jj) {
        mbar_wait(&bars->a_full[as], aphase);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t d_addr = tmem_base + (uint32_t)(jj * p.H_pad);
          const uint32_t a_addr = tmem_base + (uint32_t)(p.a_col0 + as * 32);
          umma_ts(d_addr, a_addr, bdesc, idesc, c > 0 ? 1u : 0u);
          umma_ts(d_addr, a_addr + 8, bdesc + 2, idesc, 1u);
          umma_ts(d_addr, a_addr + 16, bdesc + 4, idesc, 1u);
          umma_ts(d_addr, a_addr + 24, bdesc + 6, idesc, 1u);
          umma_commit(&bars->a_empty[as]);
        }
        __syncwarp();
        if (++as == (uint32_t)p.a_slots) { as = 0; aphase ^= 1; }
      }
      if (elect_one()) {
        // stage consumed (the producers read it before publishing A, the MMAs above read dyT): release it in every CTA of the cluster
        if (csize > 1) umma_commit_mcast(&bars->x_empty[st], cmask);
        else umma_commit(&bars->x_empty[st]);
      }
      __syncwarp();
      if (++st == (uint32_t)p.ns) { st = 0; phase ^= 1; }
    }
    if (active && n_chunks > 0) {
      if (elect_one()) umma_commit(&bars->acc_full);
      __syncwarp();
    }
  } else if (active) {
    // =============================== A producers (lane = channel i) + final epilogue ===============================
    const int q = warp & 3;
    const int il = q * 32 + lane;                                  // TMEM lane = (jsub, i): field within the accumulator, channel
    const int jsub = il / p.LW, ich = il - jsub * p.LW;
    const int irow = min(ich, p.HpQ - 1);                          // lanes past HpQ only produce rows nobody reads
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t st = 0, phase = 0, as = 0, aphase = 1;
    bool a_first = true;
    for (int c = 0; c < n_chunks; ++c) {
      mbar_wait(&bars->x_full[st], phase);
      const uint8_t* base = sStage + (size_t)st * stage_bytes;
      const uint8_t* xkrow = base + dy_bytes + (size_t)irow * 128;   // 128-byte row, 16-byte chunks XOR-swizzled by (row & 7)
      uint4 xk8[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) xk8[ch] = *reinterpret_cast<const uint4*>(xkrow + ((ch ^ (irow & 7)) << 4));
      for (int jj = 0; jj < nj; ++jj) {
        if (!a_first) {
          mbar_wait(&bars->a_empty[as], aphase);
          fence_after_sync();
        }
        // un-swizzled box of FPC rows, broadcast reads (fields >= m: TMA zero fill / zero padding rows of x0T -> a zero A tile)
        const uint8_t* x0row = base + dy_bytes + xk_bytes + (size_t)(jj * p.PACK + jsub) * 128;
        const uint32_t a_addr = tmem_base + lane_addr + (uint32_t)(p.a_col0 + as * 32);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t z[16];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int ch = hf * 4 + c4;
            const uint4 x0v = *reinterpret_cast<const uint4*>(x0row + (ch << 4));
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&xk8[ch]);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&x0v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 prod = __hmul2(a2[e], b2[e]);
              z[c4 * 4 + e] = *reinterpret_cast<uint32_t*>(&prod);
            }
          }
          tmem_st_x16(a_addr + hf * 16, z);
        }
        tmem_wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_full[as]);
        if (++as == (uint32_t)p.a_slots) { as = 0; aphase ^= 1; a_first = false; }
      }
      if (++st == (uint32_t)p.ns) { st = 0; phase ^= 1; }
    }
    // ---- epilogue: partial dW_j^T[i, :] for this r range
    if (n_chunks > 0) {
      mbar_wait(&bars->acc_full, 0);
      fence_after_sync();
    }
    // (tcgen05.ld is warp-collective: every lane runs the loads, only lanes that own a channel store)
    for (int jj = 0; jj < nj; ++jj) {
      const int field = j0 + jj * p.PACK + jsub;
      const bool own = ich < p.HpQ && field < p.m;
      float* out = p.part + (((int64_t)split * p.m + min(field, p.m - 1)) * p.HpQ + irow) * p.H_pad;
      for (int c0 = 0; c0 < p.H_pad; c0 += 16) {
        uint32_t v[16];
        if (n_chunks > 0) {
          tmem_ld_x16(tmem_base + lane_addr + (uint32_t)(jj * p.H_pad + c0), v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (own) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(out + c0 + i) =
                make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW[h, i*m + j] = sum_s part[s, j, i, h]  (fixed order), db[h] handled elsewhere.
// part is contiguous in h, dW in k = i*m + j: a 32 x 32 (k, h) tile per block goes through shared memory so that both the reads
// (32 consecutive h per (s, j, i): one 128-byte line per warp) and the writes (32 consecutive k per h) are coalesced -- the direct
// form read one 4-byte word per 32-byte sector, S times per output.
__global__ void __launch_bounds__(256) cin_dw_reduce_tc_kernel(const float* __restrict__ part, int S, int m, int Hp, int HpQ, int H, int H_pad,
                                                               float* __restrict__ dW) {
  __shared__ float tile[32][33];
  const int K = Hp * m;
  const int k0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 warps
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = k0 + ty * 4 + r, h = h0 + tx;
    float v = 0.f;
    if (k < K && h < H) {
      const int i = k / m, j = k - i * m;
      const float* src = part + ((int64_t)j * HpQ + i) * H_pad + h;
      const int64_t sstride = (int64_t)m * HpQ * H_pad;
      int s = 0;
      for (; s + 4 <= S; s += 4) {                               // four loads in flight, summed in split order
        const float a0 = src[(s + 0) * sstride], a1 = src[(s + 1) * sstride], a2 = src[(s + 2) * sstride], a3 = src[(s + 3) * sstride];
        v += a0; v += a1; v += a2; v += a3;
      }
      for (; s < S; ++s) v += src[s * sstride];
    }
    tile[ty * 4 + r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int h = h0 + ty * 4 + r, k = k0 + tx;
    if (h < H && k < K) dW[(int64_t)h * K + k] = tile[tx][ty * 4 + r];
  }
}

// ------------------------------------------------------------------------------------------------
// rows [R, pitch] bf16 (first C channels) -> channel-major [CP, R] bf16 (rows >= C zero); 64x64 tiles through shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_to_cols_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t pitch, int64_t R, int C, int CP,
                                                                __nv_bfloat16* __restrict__ dst) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];       // 144-byte rows: 16-byte aligned granules
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const bool vec = ((pitch & 7) == 0) && (((uintptr_t)src & 15) == 0);
  {
    // 16-byte granules of a source row (8 channels); channels >= C read as zero
    const int rr = threadIdx.x >> 2, cg = (threadIdx.x & 3) * 16;
    const int64_t r = r0 + rr;
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      const int c = c0 + cg + g8 * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < R && c < C) {
        if (vec && c + 8 <= C) {
          v = *reinterpret_cast<const uint4*>(src + r * pitch + c);
        } else {
          __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v);
          for (int i = 0; i < 8; ++i)
            if (c + i < C) e[i] = src[r * pitch + c + i];
        }
      }
      *reinterpret_cast<uint4*>(&tile[rr][cg + g8 * 8]) = v;
    }
  }
  __syncthreads();
  {
    // lanes walk the channel axis (conflict-free column reads), each thread writes 16 consecutive rows of one channel
    const int cc = threadIdx.x & 63, rg = (threadIdx.x >> 6) * 16;
    const int c = c0 + cc;
    if (c < CP) {
      __align__(16) __nv_bfloat16 col[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) col[i] = tile[rg + i][cc];
      const int64_t r = r0 + rg;
      __nv_bfloat16* d = dst + (int64_t)c * R + r;
      if (r + 16 <= R && ((R & 7) == 0) && (((uintptr_t)dst & 15) == 0)) {
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(&col[0]);
        *reinterpret_cast<uint4*>(d + 8) = *reinterpret_cast<const uint4*>(&col[8]);
      } else {
        for (int i = 0; i < 16; ++i)
          if (r + i < R) d[i] = col[i];
      }
    }
  }
}

extern "C" int xdfm_rows_to_cols_bf16(const void* src, int64_t pitch, int64_t R, int C, int CP, void* dst, void* stream) {
  if (R == 0 || CP == 0) return XDFM_OK;
  dim3 grid((unsigned)ceil_div64(R, 64), (unsigned)ceil_div64(CP, 64));
  rows_to_cols_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, pitch, R, C, CP, (__nv_bfloat16*)dst);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// db[h] = sum_r dyT[h, r] : one block per channel, fixed-order tree
__global__ void __launch_bounds__(256) cin_db_cols_kernel(const __nv_bfloat16* __restrict__ dyT, int64_t R, float* __restrict__ db) {
  const int h = blockIdx.x;
  const __nv_bfloat16* row = dyT + (int64_t)h * R;
  float acc = 0.f;
  if ((R & 7) == 0 && ((uintptr_t)row & 15) == 0) {
    // 16-byte granules, four in flight per thread; the order of the additions is a fixed function of R
    const uint4* row4 = reinterpret_cast<const uint4*>(row);
    const int64_t n4 = R >> 3;
    int64_t g = threadIdx.x;
    for (; g + 3 * 256 < n4; g += 4 * 256) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = row4[g + u * 256];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(b2[i]);
          acc += f.x;
          acc += f.y;
        }
      }
    }
    for (; g < n4; g += 256) {
      const uint4 v = row4[g];
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(b2[i]);
        acc += f.x;
        acc += f.y;
      }
    }
  } else {
    for (int64_t r = threadIdx.x; r < R; r += 256) acc += __bfloat162float(row[r]);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) db[h] = red[0];
}

static int round_up_w(int a, int b) { return (a + b - 1) / b * b; }
extern int g_cin_tc_cluster_shared;

// 0 = automatic (two fields per CTA whenever two accumulators + a two-slot A ring fit the 512 TMEM columns), 1 = one field per CTA
static int g_cin_dw_force_jp = 0;
extern "C" void xdfm_cin_dw_set_jp(int v) { g_cin_dw_force_jp = (v == 1) ? 1 : 0; }
// A/B switch of the lane packing (tests compare both geometries): 1 = one field per accumulator whatever the width
static int g_cin_dw_no_pack = 0;
extern "C" void xdfm_cin_dw_set_pack(int enabled) { g_cin_dw_no_pack = enabled ? 0 : 1; }

struct CinDwGeom {
  int HpQ, H_pad, mP, JP, PACK, n_jgroups, n_jgroups_padded, n_splits, ns, a_slots, a_col0, cluster;
  int64_t chunks_total, chunks_per_split;
  size_t smem;
};

static int cin_dw_geom(int64_t B, int m, int Hp, int H, int D, CinDwGeom* g) {
  if (!(D == 8 || D == 16 || D == 32 || D == 64 || D == 128) || Hp > 128 || H > 256 || m > XDFM_MAX_FIELDS) {
    xdfm_set_error("cin_bwd_dw_tc: unsupported shape (D=%d in {8..128 pow2}, Hp=%d<=128, H=%d<=256, m=%d<=64)", D, Hp, H, m);
    return XDFM_ERR_UNSUPPORTED;
  }
  g->HpQ = round_up_w(Hp, 16);
  g->H_pad = round_up_w(H, 16);
  g->mP = round_up_w(m, 8);
  g->JP = (2 * g->H_pad + 64 <= 512) ? 2 : 1;
  if (g_cin_dw_force_jp == 1) g->JP = 1;       // experiment switch (xdfm_cin_dw_set_jp): one field per CTA -> deeper A ring in TMEM
  g->PACK = g_cin_dw_no_pack ? 1 : (g->HpQ <= 32 ? 4 : (g->HpQ <= 64 ? 2 : 1));
  g->a_col0 = round_up_w(g->JP * g->H_pad, 32);
  g->a_slots = std::min(DW_MAX_ASLOTS, (512 - g->a_col0) / 32);
  g->n_jgroups = (m + g->JP * g->PACK - 1) / (g->JP * g->PACK);
  int cluster = g_cin_tc_cluster_shared;
  while (cluster > 1 && (((g->H_pad / 8) % cluster) != 0 || ((g->HpQ / 8) % cluster) != 0)) cluster >>= 1;
  g->cluster = cluster;
  g->n_jgroups_padded = round_up_w(g->n_jgroups, cluster);
  const int64_t R = B * (int64_t)D;
  g->chunks_total = ceil_div64(R, 64);
  int splits = std::max(1, xdfm_num_sms() / g->n_jgroups_padded);
  splits = (int)std::min<int64_t>(splits, std::max<int64_t>(1, g->chunks_total / 4));
  g->n_splits = splits;
  g->chunks_per_split = ceil_div64(g->chunks_total, splits);
  size_t stage = (size_t)g->H_pad * 128 + (size_t)g->HpQ * 128 + 1024;
  int ns = (int)std::min<size_t>((227 * 1024 - sizeof(CinDwBars) - 256) / stage, DW_MAX_NS);
  if (ns < 2 || g->a_slots < 2) {
    xdfm_set_error("cin_bwd_dw_tc: resources too small (ns=%d a_slots=%d)", ns, g->a_slots);
    return XDFM_ERR_UNSUPPORTED;
  }
  g->ns = ns;
  g->smem = (size_t)ns * stage + sizeof(CinDwBars) + 256;
  return XDFM_OK;
}

extern "C" int64_t xdfm_cin_bwd_dw_tc_workspace_bytes(int64_t B, int m, int Hp, int H, int D) {
  CinDwGeom g;
  if (cin_dw_geom(B, m, Hp, H, D, &g) != XDFM_OK) return -1;
  return (int64_t)g.n_splits * m * g.HpQ * g.H_pad * 4;
}

// dyT [H_pad, R] bf16, xkT [HpQ, R] bf16, x0T [mP, R] bf16 (channel-major, zero padded rows); dW fp32 [H, Hp*m] and db [H] out.
extern "C" int xdfm_cin_bwd_dw_tc(const void* dyT, const void* xkT, const void* x0T, int64_t B, int m, int Hp, int H, int D, float* dW,
                                  float* db, void* workspace, int64_t workspace_bytes, void* stream) {
  CinDwGeom g;
  int rc = cin_dw_geom(B, m, Hp, H, D, &g);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  const int64_t R = B * (int64_t)D;
  XDFM_CHECK_ARG(workspace_bytes >= (int64_t)g.n_splits * m * g.HpQ * g.H_pad * 4, "cin_bwd_dw_tc: workspace too small");
  XDFM_CHECK_ARG(R % 8 == 0, "cin_bwd_dw_tc: B*D must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmDy, tmXk, tmX0;
  rc = xdfm_make_tmap_bf16(&tmDy, dyT, (uint64_t)g.H_pad, (uint64_t)R, (uint64_t)R * 2, (uint32_t)(g.H_pad / g.cluster), 64, 1);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmXk, xkT, (uint64_t)g.HpQ, (uint64_t)R, (uint64_t)R * 2, (uint32_t)(g.HpQ / g.cluster), 64, 1);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmX0, x0T, (uint64_t)g.mP, (uint64_t)R, (uint64_t)R * 2, (uint32_t)(g.JP * g.PACK), 64, 0);
  if (rc) return rc;
  CinDwParams p;
  p.part = (float*)workspace; p.R = R; p.m = m; p.Hp = Hp; p.HpQ = g.HpQ; p.H_pad = g.H_pad; p.JP = g.JP;
  p.PACK = g.PACK; p.LW = 128 / g.PACK;
  p.n_jgroups = g.n_jgroups; p.n_jgroups_padded = g.n_jgroups_padded; p.n_splits = g.n_splits;
  p.chunks_total = g.chunks_total; p.chunks_per_split = g.chunks_per_split; p.ns = g.ns; p.a_slots = g.a_slots; p.a_col0 = g.a_col0;
  XDFM_CUDA(cudaFuncSetAttribute(cin_bwd_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.n_jgroups_padded * g.n_splits);
  cfg.blockDim = dim3(DW_THREADS);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_bwd_dw_tc_kernel, tmDy, tmXk, tmX0, p));
  XDFM_LAUNCH_CHECK();
  if (dW != nullptr) {
    dim3 rgrid((unsigned)ceil_div64((int64_t)Hp * m, 32), (unsigned)ceil_div64(H, 32));
    cin_dw_reduce_tc_kernel<<<rgrid, 256, 0, st>>>((const float*)workspace, g.n_splits, m, Hp, g.HpQ, H, g.H_pad, dW);
    XDFM_LAUNCH_CHECK();
  }
  if (db != nullptr) {
    cin_db_cols_kernel<<<H, 256, 0, st>>>((const __nv_bfloat16*)dyT, R, db);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}
