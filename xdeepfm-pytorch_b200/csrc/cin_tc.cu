// CIN layer forward on the 5th-gen tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces (reference, file:line): deepctr/layers/interaction.py:218-246 -- einsum outer product -> Conv1d(k=1) -> ReLU ->
// split-half -> sum over D.  The reference materialises Z = X^{k-1} (x) X^0 as [B, h*m, D] fp32 in HBM and runs a batched GEMM
// with N = D.  Here the contraction is ONE implicit GEMM per layer
//
//        Y[r, h] = sum_k Z[r, k] * W'[h, k],      r = (sample b, d) -> M rows,   h -> N columns,   k = (j, i) -> K
//
// whose A operand Z never exists in memory: 128 producer threads (one per row r) hold X^{k-1}[r, :] in registers, multiply by
// X^0[r, j] and write packed bf16 pairs straight into TENSOR MEMORY with tcgen05.st; tcgen05.mma (kind::f16, M=128,
// N=H_pad<=256, K=16) reads A from TMEM and B = W' (bf16, K-major, SWIZZLE_128B) from shared memory where TMA put it; the fp32
// accumulator [128 x H_pad] lives in TMEM and the epilogue (tcgen05.ld -> bias -> ReLU -> split/pool/store) reads it back.
//
// Activations of this path use a ROW layout: one row per (sample, d), channels contiguous --
//   x0t [B*D, mP]  bf16 (mP = m rounded up to 8, zero padded),   y_k [B*D, Hs_k] bf16 (Hs = H rounded up to 8),
// so a producer thread reads its operand row with 128-bit loads, the epilogue writes 16 channels with two 128-bit stores, and
// one accumulator tile = 128 consecutive rows.
//
// K ordering is private to this file: k' = j * HpP + i (j over X^0 fields, i over X^{k-1} channels padded to a multiple of 8),
// W' = cin_prep_w(W) is the reference weight permuted/padded/converted accordingly each step.
//
// W' (up to 1.4 MB) is streamed once per tile; to keep that off the L2 -> SM path the CTAs of a thread-block cluster walk the
// chunk sequence in lock-step and every chunk is fetched ONCE per cluster: CTA c loads rows slice c and TMA-multicasts it into
// the shared memory of all CTAs of the cluster.
//
// TMEM columns (512 allocated): [0, 256) accumulator, [256, 512) a ring of 4 A slots (128 K-values = 64 columns each) that is
// aligned with the 128-wide W' stages (two 64-wide TMA boxes per barrier), so the MMA warp's inner loop is "wait two barriers,
// issue eight MMAs, commit": the issue loop costs ~500 cycles per iteration (ncu, round 1), eight MMAs are 832 cycles of tensor work.
// Warp roles (14 warps): 0 = TMA (W' chunks, x tiles), 1 = MMA issuer + TMEM alloc, 2..5 = Z producers, 6..13 = epilogue
// (two warps per TMEM lane quarter, alternating 16-column chunks).
#include "tc_common.cuh"
#include "../../include/xdfm.h"

using namespace tc;

#define TC_THREADS 448
#define TC_A_COL0 256
#define TC_A_SLOTS 4        // A ring slots
#define TC_A_SLOT_COLS 64  // 128 bf16 K-values = one W' stage (two 64-wide TMA boxes): 8 MMAs per barrier round trip
#define TC_MAX_NS_W 4      // W' stage ring depth (run-time, limited by shared memory)

struct CinTcParams {
  const __nv_bfloat16* x0t;   // [R, mP]
  const float* bias;          // [H]
  __nv_bfloat16* yt;          // [R, Hs] post-activation (next layer's input / backward)
  float* pooled;              // [B, fm_total] or null
  float* maps;                // [B, fm_total, D] or null
  int64_t R;                  // rows = B * D
  int m, mP, Hp, H, H_pad, Hs, D, act, hdb, fm_total, col_off;
  int64_t n_tiles;
  int n_iters;                // tile iterations per CTA (uniform over the grid)
  int n_wchunks;              // 128-wide K stages per tile (K' = m*HpP padded to a multiple of 128)
  int ns_w;                   // W' ring depth (stages)
};

struct __align__(8) CinTcBars {
  uint64_t w_full[TC_MAX_NS_W], w_empty[TC_MAX_NS_W];
  uint64_t a_full[TC_A_SLOTS], a_empty[TC_A_SLOTS];
  uint64_t x_full[2], x_empty[2];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

// ---- Z producer helpers -------------------------------------------------------------------------------------------------
struct ARing {
  CinTcBars* bars;
  uint32_t base;      // TMEM address of ring column 0 for this warp's lane quarter
  uint32_t as;        // slot being filled
  uint32_t aphase;    // parity of the "slot free" phase to wait for
  bool first;         // first pass over the ring: slots are free, nothing to wait for
  int lane;
};

__device__ __forceinline__ void ring_acquire(ARing& r) {   // entering slot r.as at column 0
  if (!r.first) {
    mbar_wait(&r.bars->a_empty[r.as], r.aphase);
    fence_after_sync();
  }
}
__device__ __forceinline__ void ring_publish(ARing& r) {   // slot r.as completely written by this warp
  tmem_wait_st();
  fence_before_sync();
  __syncwarp();
  if (r.lane == 0) mbar_arrive(&r.bars->a_full[r.as]);
  if (++r.as == TC_A_SLOTS) { r.as = 0; r.aphase ^= 1; r.first = false; }
}

// products xk2[OFF .. OFF+N) * xv2 -> N TMEM columns at addr (N = 4, 8 or 16)
template <int NPAIR, int OFF, int N>
__device__ __forceinline__ void st_products(const __nv_bfloat162 (&xk2)[NPAIR], __nv_bfloat162 xv2, uint32_t addr) {
  uint32_t z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    __nv_bfloat162 prod = __hmul2(xk2[OFF + i], xv2);
    z[i] = *reinterpret_cast<uint32_t*>(&prod);
  }
  if constexpr (N == 16) tmem_st_x16(addr, z);
  else if constexpr (N == 8) tmem_st_x8(addr, z);
  else tmem_st_x4(addr, z);
}
template <int NPAIR, int OFF, int SEG>   // SEG columns (multiple of 4) in the widest shapes
__device__ __forceinline__ void st_segment(const __nv_bfloat162 (&xk2)[NPAIR], __nv_bfloat162 xv2, uint32_t addr) {
  if constexpr (SEG > 0) {
    constexpr int N = SEG >= 16 ? 16 : (SEG >= 8 ? 8 : 4);
    st_products<NPAIR, OFF, N>(xk2, xv2, addr);
    st_segment<NPAIR, OFF + N, SEG - N>(xk2, xv2, addr + N);
  }
}
// One X^0 field j: NPAIR columns starting at column CIN of the current slot; slot boundaries are compile-time positions.
template <int NPAIR, int CIN, int OFF>
__device__ __forceinline__ void emit_field(const __nv_bfloat162 (&xk2)[NPAIR], __nv_bfloat162 xv2, ARing& r) {
  if constexpr (OFF < NPAIR) {
    if constexpr (CIN == 0) ring_acquire(r);
    constexpr int room = TC_A_SLOT_COLS - CIN;
    constexpr int rem = NPAIR - OFF;
    constexpr int seg = rem < room ? rem : room;
    st_segment<NPAIR, OFF, seg>(xk2, xv2, r.base + r.as * TC_A_SLOT_COLS + CIN);
    if constexpr (CIN + seg == TC_A_SLOT_COLS) {
      ring_publish(r);
      emit_field<NPAIR, 0, OFF + seg>(xk2, xv2, r);
    }
  }
}

struct EpiRow {
  __nv_bfloat16* yt;   // &yt[r, 0] or null
  float* maps;         // &maps[b, col_off - hdb, d] or null (index by h, stride D)
  float* pooled;       // &pooled[b, col_off - hdb] or null (index by h)
  float* spool;        // per-lane-quarter scratch [H_pad] (D > 32)
  float act_floor;     // 0 for ReLU, -inf for linear
  bool valid;
};

// One 16-column accumulator chunk of one row: bias + activation, bf16 / fp32 stores, and the sum over the sample's D lanes.
// The lane-sum uses a transposing butterfly: every shuffle level halves the number of live values per lane, so 16 columns
// cost 8+4+2+1 shuffles instead of 16*log2(LD); afterwards each lane holds the finished sum of its own column(s).
// Rows past the end of the batch need no masking: a sample's D rows are all valid or all invalid, so garbage never mixes
// into a valid sample, and invalid rows have null output pointers.  Columns >= H are exact zeros for ReLU / linear
// (zero rows of W', zero bias) and are masked explicitly for sigmoid.
template <int LD, int ACT>   // LD = lanes per sample inside the warp: min(D, 32)
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[16], int c0, int lane, const float* __restrict__ sBias,
                                               const CinTcParams& p, const EpiRow& r) {
  float y[16];
#pragma unroll
  for (int g4 = 0; g4 < 4; ++g4) {
    const float4 bv = *reinterpret_cast<const float4*>(sBias + c0 + g4 * 4);
    y[g4 * 4 + 0] = __uint_as_float(v[g4 * 4 + 0]) + bv.x;
    y[g4 * 4 + 1] = __uint_as_float(v[g4 * 4 + 1]) + bv.y;
    y[g4 * 4 + 2] = __uint_as_float(v[g4 * 4 + 2]) + bv.z;
    y[g4 * 4 + 3] = __uint_as_float(v[g4 * 4 + 3]) + bv.w;
  }
  if constexpr (ACT != XDFM_ACT_SIGMOID) {
    // ReLU and linear share one path: max(y, floor) with floor = 0 or -inf
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], r.act_floor);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = (c0 + i < p.H) ? 1.f / (1.f + __expf(-y[i])) : 0.f;
  }
  if (r.yt != nullptr) {
    // channels [c0, c0+16) of this row; columns H..Hs-1 are zeros, columns >= Hs do not exist
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      if (c0 + g8 * 8 < p.Hs) {
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(y[g8 * 8 + 0], y[g8 * 8 + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(y[g8 * 8 + 2], y[g8 * 8 + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(y[g8 * 8 + 4], y[g8 * 8 + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(y[g8 * 8 + 6], y[g8 * 8 + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(r.yt + c0 + g8 * 8) = pk;
      }
    }
  }
  if (r.maps != nullptr) {
    float* mp = r.maps + (int64_t)c0 * p.D;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < p.H && c0 + i >= p.hdb) mp[(int64_t)i * p.D] = y[i];
  }
  if (p.pooled == nullptr || c0 + 16 <= p.hdb) return;    // no direct channel in this chunk
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

// All 16-column chunks of one accumulator tile that belong to this warp (c0 = 16*half, +32, ...); the tcgen05.ld of the next
// chunk is in flight while the current one is processed.
template <int LD, int ACT>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_row, int half, int lane, const float* __restrict__ sBias,
                                              const CinTcParams& p, const EpiRow& r) {
  uint32_t va[16], vb[16];
  int c0 = half * 16;
  if (c0 < p.H_pad) tmem_ld_x16(tmem_row + c0, va);
  for (; c0 < p.H_pad; c0 += 64) {
    tmem_wait_ld();
    if (c0 + 32 < p.H_pad) tmem_ld_x16(tmem_row + c0 + 32, vb);
    epilogue_chunk<LD, ACT>(va, c0, lane, sBias, p, r);
    if (c0 + 32 < p.H_pad) {
      tmem_wait_ld();
      if (c0 + 64 < p.H_pad) tmem_ld_x16(tmem_row + c0 + 64, va);
      epilogue_chunk<LD, ACT>(vb, c0 + 32, lane, sBias, p, r);
    }
  }
}

template <int LD>
__device__ __forceinline__ void epilogue_tile_act(uint32_t tmem_row, int half, int lane, const float* __restrict__ sBias,
                                                  const CinTcParams& p, const EpiRow& r) {
  epilogue_tile<LD, XDFM_ACT_RELU>(tmem_row, half, lane, sBias, p, r);
}

// NI8 = HpP / 8 (channels of X^{k-1} padded to a multiple of 8, HpP <= 128)
template <int NI8>
__global__ void __launch_bounds__(TC_THREADS, 1)
cin_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmXk, CinTcParams p) {
  constexpr int HpP = NI8 * 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- shared memory carve-up
  const uint32_t w_box_bytes = (uint32_t)p.H_pad * 128;                     // one 64-wide K chunk
  const uint32_t w_stage_bytes = 2 * w_box_bytes;
  uint8_t* sW = smem;                                                       // ns_w x 2 x [H_pad x 128 B], 1024-aligned
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;                        // [128][mP] bf16
  const uint32_t xk_tile = (uint32_t)128 * HpP * 2;                         // [128][HpP] bf16
  uint8_t* sX0 = sW + (size_t)p.ns_w * w_stage_bytes;                       // 2 tiles
  uint8_t* sXk = sX0 + 2 * (size_t)x0_tile;                                 // 2 tiles
  float* sBias = reinterpret_cast<float*>(sXk + 2 * (size_t)xk_tile);       // [H_pad]
  float* sPool = sBias + p.H_pad;                                           // [4][H_pad] pooling scratch (D > 32)
  CinTcBars* bars = reinterpret_cast<CinTcBars*>(sPool + 4 * p.H_pad);

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_MAX_NS_W; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < TC_A_SLOTS; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], 4); }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, 8);
    fence_barrier_init();
  }
  for (int h = threadIdx.x; h < p.H_pad; h += TC_THREADS) sBias[h] = h < p.H ? p.bias[h] : 0.f;
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();            // remote barriers must be initialised before any multicast / remote arrive
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  // every CTA runs n_iters iterations so that the CTAs of a cluster stay in lock-step on the W' chunk stream;
  // iterations whose tile index is out of range only keep the W' ring moving.
  auto tile_of = [&](int it) -> int64_t { return (int64_t)it * gridDim.x + blockIdx.x; };

  if (warp == 0) {
    // =============================== TMA producer: x tiles (one tile ahead) and the W' chunk stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmXk);
      // W' rows [wr0, wr1) are this CTA's multicast slice (the host picks a cluster size that divides H_pad/8)
      const int slice = p.H_pad / (int)csize;
      const int wr0 = (int)crank * slice;
      int xit = 0;                               // number of x tiles loaded so far
      auto load_x = [&](int64_t tile) {
        const int buf = xit & 1;
        if (xit >= 2) mbar_wait(&bars->x_empty[buf], ((xit >> 1) - 1) & 1);
        const int64_t r0 = tile * 128;
        const uint32_t nrows = (uint32_t)min((int64_t)128, p.R - r0);
        mbar_arrive_expect_tx(&bars->x_full[buf], nrows * (uint32_t)(p.mP * 2) + xk_tile);
        bulk_load_1d(sX0 + (size_t)buf * x0_tile, p.x0t + r0 * p.mP, nrows * (uint32_t)(p.mP * 2), &bars->x_full[buf]);
        tma_load_2d(sXk + (size_t)buf * xk_tile, &tmXk, 0, (int)r0, &bars->x_full[buf]);   // OOB rows are zero-filled
        ++xit;
      };
      uint32_t ws = 0, wphase = 1;  // ring slot; parity of the "slot free" phase to wait for (first pass: free already)
      bool first_pass = true;
      if (tile_of(0) < p.n_tiles) load_x(tile_of(0));
      for (int it = 0; it < p.n_iters; ++it) {
        if (it + 1 < p.n_iters && tile_of(it + 1) < p.n_tiles) load_x(tile_of(it + 1));
        for (int c = 0; c < p.n_wchunks; ++c) {
          if (!first_pass) mbar_wait(&bars->w_empty[ws], wphase);
          mbar_arrive_expect_tx(&bars->w_full[ws], w_stage_bytes);
          uint8_t* dst = sW + (size_t)ws * w_stage_bytes + (size_t)wr0 * 128;
          if (csize > 1) {
            tma_load_2d_mcast(dst, &tmW, c * 128, wr0, &bars->w_full[ws], cmask);
            tma_load_2d_mcast(dst + w_box_bytes, &tmW, c * 128 + 64, wr0, &bars->w_full[ws], cmask);
          } else {
            tma_load_2d(dst, &tmW, c * 128, wr0, &bars->w_full[ws]);
            tma_load_2d(dst + w_box_bytes, &tmW, c * 128 + 64, wr0, &bars->w_full[ws]);
          }
          if (++ws == (uint32_t)p.ns_w) { ws = 0; wphase ^= 1; first_pass = false; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the loop (warp-uniform control flow and operands stay in uniform registers); one elected lane
    // issues the tcgen05 instructions.  Per 64-wide K chunk: wait W' + A, issue four K=16 MMAs, release both slots.
    const uint32_t idesc = make_idesc_bf16(128, p.H_pad);
    const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
    const uint32_t stage_desc_step = w_stage_bytes >> 4;       // descriptor start-address units (16 B)
    const uint32_t box_desc_step = w_box_bytes >> 4;
    uint32_t ws = 0, wphase = 0;                                // W' ring slot / phase parity
    uint64_t bdesc = bdesc0;                                    // descriptor of ring slot ws
    uint32_t as = 0, aphase = 0;                                // A ring slot / phase parity
    int at = 0;                                                 // active tiles so far
    for (int it = 0; it < p.n_iters; ++it) {
      const bool active = tile_of(it) < p.n_tiles;
      if (active && at > 0) {
        mbar_wait(&bars->acc_empty, (at - 1) & 1);
        fence_after_sync();
      }
      for (int c = 0; c < p.n_wchunks; ++c) {
        mbar_wait(&bars->w_full[ws], wphase);
        if (active) mbar_wait(&bars->a_full[as], aphase);
        fence_after_sync();
        if (elect_one()) {
          if (active) {
            const uint32_t a_addr = tmem_base + TC_A_COL0 + as * TC_A_SLOT_COLS;
            umma_ts(tmem_base, a_addr, bdesc, idesc, c > 0 ? 1u : 0u);
            umma_ts(tmem_base, a_addr + 8, bdesc + 2, idesc, 1u);
            umma_ts(tmem_base, a_addr + 16, bdesc + 4, idesc, 1u);
            umma_ts(tmem_base, a_addr + 24, bdesc + 6, idesc, 1u);
            const uint64_t bdesc1 = bdesc + box_desc_step;
            umma_ts(tmem_base, a_addr + 32, bdesc1, idesc, 1u);
            umma_ts(tmem_base, a_addr + 40, bdesc1 + 2, idesc, 1u);
            umma_ts(tmem_base, a_addr + 48, bdesc1 + 4, idesc, 1u);
            umma_ts(tmem_base, a_addr + 56, bdesc1 + 6, idesc, 1u);
            umma_commit(&bars->a_empty[as]);     // A chunk consumed
          }
          // W' chunk consumed by this CTA: tell every CTA of the cluster (each may overwrite this slot by multicast)
          if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
          else umma_commit(&bars->w_empty[ws]);
        }
        __syncwarp();
        if (++ws == (uint32_t)p.ns_w) { ws = 0; wphase ^= 1; bdesc = bdesc0; }
        else bdesc += stage_desc_step;
        if (active && ++as == TC_A_SLOTS) { as = 0; aphase ^= 1; }
      }
      if (active) {
        if (elect_one()) umma_commit(&bars->acc_full);   // accumulator complete
        __syncwarp();
        ++at;
      }
    }
  } else if (warp < 6) {
    // =============================== Z producers: thread <-> accumulator row ===============================
    // Z[r, k'] for k' = j*HpP + i is written in k' order into the A ring, one 8-value granule (4 columns) at a time; a ring slot
    // (64 K-values) is published as soon as its last granule has been stored.
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int rl = q * 32 + lane;                // row within the tile
    ARing ring;
    ring.bars = bars;
    ring.base = tmem_base + ((uint32_t)(q * 32) << 16) + TC_A_COL0;
    ring.as = 0; ring.aphase = 1; ring.first = true; ring.lane = lane;
    int at = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      if (tile_of(it) >= p.n_tiles) continue;
      const int buf = at & 1;
      mbar_wait(&bars->x_full[buf], (at >> 1) & 1);
      const __nv_bfloat16* x0row = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)buf * x0_tile) + (size_t)rl * p.mP;
      const uint4* xkrow = reinterpret_cast<const uint4*>(sXk + (size_t)buf * xk_tile + (size_t)rl * HpP * 2);
      // X^{k-1}[r, 0..HpP) as packed bf16 pairs: channels >= Hp are other data / padding and meet zero columns of W'
      __nv_bfloat162 xk2[HpP / 2];
#pragma unroll
      for (int v8 = 0; v8 < NI8; ++v8) {
        const uint4 t = xkrow[v8];
        xk2[v8 * 4 + 0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        xk2[v8 * 4 + 1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        xk2[v8 * 4 + 2] = *reinterpret_cast<const __nv_bfloat162*>(&t.z);
        xk2[v8 * 4 + 3] = *reinterpret_cast<const __nv_bfloat162*>(&t.w);
      }
      int ph = 0;                                // granule (4 columns) inside the current ring slot where the next field starts
      for (int j = 0; j < p.m; ++j) {
        const __nv_bfloat16 xv = x0row[j];
        const __nv_bfloat162 xv2 = __halves2bfloat162(xv, xv);
        switch (ph) {
          case 0: emit_field<HpP / 2, 0, 0>(xk2, xv2, ring); break;
          case 1: emit_field<HpP / 2, 4, 0>(xk2, xv2, ring); break;
          case 2: emit_field<HpP / 2, 8, 0>(xk2, xv2, ring); break;
          case 3: emit_field<HpP / 2, 12, 0>(xk2, xv2, ring); break;
          case 4: emit_field<HpP / 2, 16, 0>(xk2, xv2, ring); break;
          case 5: emit_field<HpP / 2, 20, 0>(xk2, xv2, ring); break;
          case 6: emit_field<HpP / 2, 24, 0>(xk2, xv2, ring); break;
          case 7: emit_field<HpP / 2, 28, 0>(xk2, xv2, ring); break;
          case 8: emit_field<HpP / 2, 32, 0>(xk2, xv2, ring); break;
          case 9: emit_field<HpP / 2, 36, 0>(xk2, xv2, ring); break;
          case 10: emit_field<HpP / 2, 40, 0>(xk2, xv2, ring); break;
          case 11: emit_field<HpP / 2, 44, 0>(xk2, xv2, ring); break;
          case 12: emit_field<HpP / 2, 48, 0>(xk2, xv2, ring); break;
          case 13: emit_field<HpP / 2, 52, 0>(xk2, xv2, ring); break;
          case 14: emit_field<HpP / 2, 56, 0>(xk2, xv2, ring); break;
          default: emit_field<HpP / 2, 60, 0>(xk2, xv2, ring); break;
        }
        ph = (ph + NI8) & 15;
      }
      if (ph != 0) {                             // zero-fill the tail of the last K stage and publish it
        const uint32_t zz[4] = {0u, 0u, 0u, 0u};
        for (; ph < 16; ++ph) tmem_st_x4(ring.base + ring.as * TC_A_SLOT_COLS + ph * 4, zz);
        ring_publish(ring);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_empty[buf]);
      ++at;
    }
  } else {
    // =============================== epilogue: TMEM -> bias -> activation -> y / pooled / maps ===============================
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int half = (warp - 6) >> 2;            // which of the two warps of this lane quarter: chunks c0 = 16*half, +32, ...
    int at = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      const int64_t tile = tile_of(it);
      if (tile >= p.n_tiles) continue;
      const int64_t row = tile * 128 + rl;
      const int64_t b = row / p.D;
      const int d = (int)(row - b * p.D);
      EpiRow r;
      r.valid = row < p.R;
      r.yt = (p.yt && r.valid) ? p.yt + row * p.Hs : nullptr;
      r.maps = (p.maps && r.valid) ? p.maps + (b * p.fm_total + p.col_off - p.hdb) * (int64_t)p.D + d : nullptr;
      r.pooled = (p.pooled && r.valid) ? p.pooled + b * p.fm_total + p.col_off - p.hdb : nullptr;
      r.spool = sPool + q * p.H_pad;
      mbar_wait(&bars->acc_full, at & 1);
      fence_after_sync();
      r.act_floor = p.act == XDFM_ACT_RELU ? 0.f : __int_as_float(0xff800000);
      if (p.D == 8) epilogue_tile_act<8>(tmem_base + lane_addr, half, lane, sBias, p, r);
      else if (p.D == 16) epilogue_tile_act<16>(tmem_base + lane_addr, half, lane, sBias, p, r);
      else epilogue_tile_act<32>(tmem_base + lane_addr, half, lane, sBias, p, r);      // D = 32, 64, 128: full-warp segments
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);   // accumulator drained: the next tile's MMAs may start
      if (p.pooled && p.D > 32) {
        // D = 64: lane quarters (0,1) and (2,3) hold one sample each; D = 128: all four quarters hold one sample
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int qps = p.D / 32;                // lane quarters per sample
        for (int e = threadIdx.x - 6 * 32; e < (4 / qps) * p.H_pad; e += 256) {
          const int sidx = e / p.H_pad, h = e - sidx * p.H_pad;
          const int64_t bb = (tile * 128) / p.D + sidx;
          if (bb * p.D < p.R && h >= p.hdb && h < p.H) {
            float s = 0.f;
            for (int w2 = 0; w2 < qps; ++w2) s += sPool[(sidx * qps + w2) * p.H_pad + h];
            p.pooled[bb * p.fm_total + p.col_off + (h - p.hdb)] = s;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      ++at;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();            // no CTA may exit while a peer can still multicast into it / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

#include "cin_tc_dual.cuh"

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

// x [B, C, D] fp32 -> xt [B*D, CP] bf16 (row layout, channels contiguous, zero padded to CP)
__global__ void to_rows_bf16_kernel(const float* __restrict__ x, int64_t B, int C, int D, int CP, __nv_bfloat16* __restrict__ xt) {
  int64_t total = B * (int64_t)D * CP;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % CP);
    int64_t r = e / CP;
    int d = (int)(r % D);
    int64_t b = r / D;
    xt[e] = __float2bfloat16(c < C ? x[(b * C + c) * (int64_t)D + d] : 0.f);
  }
}

extern "C" int xdfm_to_rows_bf16(const float* x, int64_t B, int C, int D, int CP, void* xt, void* stream) {
  XDFM_CHECK_ARG(CP >= C && CP % 8 == 0, "to_rows_bf16: CP=%d must be a multiple of 8 and >= C=%d", CP, C);
  int64_t total = B * (int64_t)D * CP;
  if (total == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  to_rows_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, B, C, D, CP, (__nv_bfloat16*)xt);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct CinTcGeom {
  int HpP, H_pad, Hs, mP, KP, n_wchunks, ns_w;
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
  g->Hs = round_up(H, 8);
  g->mP = round_up(m, 8);
  g->n_wchunks = (m * g->HpP + 127) / 128;     // 128-wide stages
  g->KP = g->n_wchunks * 128;
  size_t fixed = 2 * (size_t)128 * g->mP * 2 + 2 * (size_t)128 * g->HpP * 2 + (size_t)5 * g->H_pad * 4 + sizeof(CinTcBars) + 256;
  size_t stage = (size_t)g->H_pad * 256;
  int ns = (int)((227 * 1024 - fixed) / stage);
  ns = std::min(ns, TC_MAX_NS_W);
  if (ns < 2) {
    xdfm_set_error("cin_tc: shared memory too small for m=%d Hp=%d H=%d D=%d", m, Hp, H, D);
    return XDFM_ERR_UNSUPPORTED;
  }
  g->ns_w = ns;
  g->smem = fixed + (size_t)ns * stage;
  return XDFM_OK;
}

extern "C" int64_t xdfm_cin_tc_wprime_elems(int m, int Hp, int H, int D) {
  CinTcGeom g;
  if (cin_tc_geom(m, Hp, H, D, &g) != XDFM_OK) return -1;
  return (int64_t)g.H_pad * g.KP;
}

int g_cin_tc_cluster_shared = 2;   // cluster size of the weight-stream multicast (all tensor-core CIN kernels)
extern "C" void xdfm_cin_tc_set_cluster(int c) { g_cin_tc_cluster_shared = (c == 1 || c == 2 || c == 4) ? c : 2; }

template <int NI8>
static int launch_cin_fwd_tc(const CUtensorMap& tmW, const CUtensorMap& tmXk, const CinTcParams& p, size_t smem, int blocks, int cluster,
                             cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_fwd_tc_kernel<NI8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_fwd_tc_kernel<NI8>, tmW, tmXk, p));
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

template <int NI8>
static int launch_cin_fwd_tc_dual(const CUtensorMap& tmW, const CUtensorMap& tmXk, const CinTcParams& p, size_t smem, int blocks, int cluster,
                                  cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_fwd_tc_dual_kernel<NI8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(TP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_fwd_tc_dual_kernel<NI8>, tmW, tmXk, p));
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

int g_cin_tc_pair = 1;      // 1: dual-producer forward kernel whenever the shape allows it (default), 0: original warp-specialised kernel
extern "C" void xdfm_cin_tc_set_pair(int v) { g_cin_tc_pair = v ? 1 : 0; }

// Row-layout operands: x0t [B*D, mP] bf16 (mP = m rounded up to 8); xkt = layer input rows with pitch xk_pitch (elements), the
// layer uses its first Hp channels; W fp32 [H, Hp*m], bias [H]; wprime = bf16 scratch [xdfm_cin_tc_wprime_elems].
// yt [B*D, Hs] bf16 out (Hs = H rounded up to 8); pooled / maps as in xdfm_cin_fwd_f32.
extern "C" int xdfm_cin_fwd_tc(const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, const float* bias, void* wprime,
                               int64_t B, int m, int Hp, int H, int D, int act, void* yt, int direct_begin, float* pooled, float* maps,
                               int fm_total, int col_off, void* stream) {
  CinTcGeom g;
  int rc = cin_tc_geom(m, Hp, H, D, &g);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  if (act != XDFM_ACT_RELU && act != XDFM_ACT_NONE) {
    xdfm_set_error("cin_fwd_tc: activation %d is not fused in the bf16 tensor-core path (relu / linear); use cin_precision='fp32'", act);
    return XDFM_ERR_UNSUPPORTED;
  }
  XDFM_CHECK_ARG(((uintptr_t)x0t % 16 == 0) && ((uintptr_t)xkt % 16 == 0) && ((uintptr_t)yt % 16 == 0) && xk_pitch % 8 == 0 &&
                     xk_pitch >= g.HpP,
                 "cin_fwd_tc: operands must be 16-byte aligned, xk_pitch (%lld) a multiple of 8 and >= %d", (long long)xk_pitch, g.HpP);
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t total = (int64_t)g.H_pad * g.KP;
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64(total, 256));
    cin_prep_w_kernel<<<blocks, 256, 0, st>>>(W, H, Hp, m, g.HpP, g.H_pad, g.KP, (__nv_bfloat16*)wprime);
    XDFM_LAUNCH_CHECK();
  }
  const int64_t R = B * (int64_t)D;
  CUtensorMap tmW, tmXk;
  int cluster = g_cin_tc_cluster_shared;
  while (cluster > 1 && ((g.H_pad / 8) % cluster) != 0) cluster >>= 1;
  // one TMA box = one CTA's multicast slice of a 64-wide W' chunk
  rc = xdfm_make_tmap_bf16(&tmW, wprime, (uint64_t)g.H_pad, (uint64_t)g.KP, (uint64_t)g.KP * 2, (uint32_t)(g.H_pad / cluster), 64, 1);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmXk, xkt, (uint64_t)R, (uint64_t)xk_pitch, (uint64_t)xk_pitch * 2, 128, (uint32_t)g.HpP, 0);
  if (rc) return rc;
  CinTcParams p;
  p.x0t = (const __nv_bfloat16*)x0t; p.bias = bias; p.yt = (__nv_bfloat16*)yt; p.pooled = pooled; p.maps = maps; p.R = R;
  p.m = m; p.mP = g.mP; p.Hp = Hp; p.H = H; p.H_pad = g.H_pad; p.Hs = g.Hs; p.D = D; p.act = act; p.hdb = direct_begin;
  p.fm_total = fm_total; p.col_off = col_off;
  p.n_tiles = ceil_div64(R, 128); p.n_wchunks = g.n_wchunks;
  p.ns_w = g.ns_w;
  int sms = xdfm_num_sms();
  int blocks = (int)std::min<int64_t>(ceil_div64(p.n_tiles, cluster) * cluster, (int64_t)(sms / cluster) * cluster);
  blocks = std::max(blocks, cluster);
  p.n_iters = (int)ceil_div64(p.n_tiles, blocks);
  if (g_cin_tc_pair && D <= 32) {
    // dual-producer kernel (cin_tc_dual.cuh): same tile schedule and W' layout, 128-wide K stages
    const size_t fixed = 2 * (size_t)128 * g.mP * 2 + 2 * (size_t)128 * g.HpP * 2 + (size_t)g.H_pad * 4 + sizeof(CinTpBars) + 256;
    const size_t stage = (size_t)g.H_pad * 256;
    int ns = (int)std::min<size_t>((227 * 1024 - fixed) / stage, TP_MAX_NS_W);
    if (ns >= 2) {
      p.ns_w = ns;
      const size_t smem_dual = fixed + (size_t)ns * stage;
      switch (g.HpP / 8) {
#define CASE_NI8P(n) case n: return launch_cin_fwd_tc_dual<n>(tmW, tmXk, p, smem_dual, blocks, cluster, st);
        CASE_NI8P(1) CASE_NI8P(2) CASE_NI8P(3) CASE_NI8P(4) CASE_NI8P(5) CASE_NI8P(6) CASE_NI8P(7) CASE_NI8P(8)
        CASE_NI8P(9) CASE_NI8P(10) CASE_NI8P(11) CASE_NI8P(12) CASE_NI8P(13) CASE_NI8P(14) CASE_NI8P(15) CASE_NI8P(16)
#undef CASE_NI8P
      }
    }
  }
  switch (g.HpP / 8) {
#define CASE_NI8(n) case n: return launch_cin_fwd_tc<n>(tmW, tmXk, p, g.smem, blocks, cluster, st);
    CASE_NI8(1) CASE_NI8(2) CASE_NI8(3) CASE_NI8(4) CASE_NI8(5) CASE_NI8(6) CASE_NI8(7) CASE_NI8(8)
    CASE_NI8(9) CASE_NI8(10) CASE_NI8(11) CASE_NI8(12) CASE_NI8(13) CASE_NI8(14) CASE_NI8(15) CASE_NI8(16)
#undef CASE_NI8
  }
  xdfm_set_error("cin_fwd_tc: unreachable HpP=%d", g.HpP);
  return XDFM_ERR_UNSUPPORTED;
}
