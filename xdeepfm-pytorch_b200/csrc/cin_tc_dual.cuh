// CIN forward, "dual producer" variant of the single-tile kernel in cin_tc.cu.
//
// Why: ncu + switch-off experiments of round 1 (profiles/r01_cin_findings.md) showed the tensor pipe only ~42 % busy with the MMA
// warp neither starved by the W' ring nor blocked on issue: the Z producers were latency-bound.  A producer warp may only touch
// its own TMEM lane quarter, and every ring slot costs it a full HMUL2 -> tcgen05.st -> tcgen05.wait::st -> mbarrier round trip
// (~900 cycles) that nothing overlapped, while the eight epilogue warps idled until the tile's last MMA.  Here the same eight
// "row" warps (two per lane quarter) do both jobs: while the MMAs run they generate Z, warp parity 0 the even ring slots and
// parity 1 the odd ones (two slots in flight per quarter), and when the accumulator is complete they drain it exactly like the
// old epilogue warps (parity = column half).  A first attempt that instead kept one producer warp per quarter and contracted TWO
// tiles per streamed W' chunk was correct but slower (0.265 vs 0.159 ms, cfg2 layer 2): the weight stream is not the limit.
//
// TMEM: [0, 256) accumulator, [256, 512) ring of 4 A slots of 64 columns (128 K-values = one W' stage of two 64-wide TMA boxes).
// Warps (10): 0 = TMA, 1 = MMA issuer + TMEM alloc, 2..9 = row warps (warp & 3 = lane quarter, (warp - 2) >> 2 = parity).
// Included by cin_tc.cu (uses its Z-producer / epilogue helpers).  D <= 32 only (the D > 32 pooling needs a cross-warp pass).
#pragma once

#define TP_THREADS 320
#define TP_SLOT_COLS 64     // 128 bf16 K-values
#define TP_ASLOTS 4
#define TP_A_COL0 256
#define TP_MAX_NS_W 4       // W' stage ring depth (stage = two 64-wide boxes)

struct __align__(8) CinTpBars {
  uint64_t w_full[TP_MAX_NS_W], w_empty[TP_MAX_NS_W];
  uint64_t a_full[TP_ASLOTS], a_empty[TP_ASLOTS];
  uint64_t x_full[2], x_empty[2];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

struct ARing2 {
  CinTpBars* bars;
  uint32_t base;      // TMEM address of ring column 0 for this warp's lane quarter
  uint32_t g;         // global sequence number of the slot being passed (all tiles of this CTA)
  uint32_t parity;    // this warp fills the slots with (g & 1) == parity
  uint32_t as;        // g % TP_ASLOTS
  bool mine;
  int lane;
};

__device__ __forceinline__ void ring2_begin(ARing2& r) {    // entering slot g at column 0
  r.as = r.g & (TP_ASLOTS - 1);
  r.mine = (r.g & 1u) == r.parity;
  if (r.mine && r.g >= TP_ASLOTS) {
    mbar_wait(&r.bars->a_empty[r.as], ((r.g / TP_ASLOTS) - 1) & 1);
    fence_after_sync();
  }
}
__device__ __forceinline__ void ring2_end(ARing2& r) {      // slot g completely passed
  if (r.mine) {
    tmem_wait_st();
    fence_before_sync();
    __syncwarp();
    if (r.lane == 0) mbar_arrive(&r.bars->a_full[r.as]);
  }
  ++r.g;
}

template <int NPAIR, int CIN, int OFF>
__device__ __forceinline__ void emit_field2(const __nv_bfloat162 (&xk2)[NPAIR], __nv_bfloat162 xv2, ARing2& r) {
  if constexpr (OFF < NPAIR) {
    if constexpr (CIN == 0) ring2_begin(r);
    constexpr int room = TP_SLOT_COLS - CIN;
    constexpr int rem = NPAIR - OFF;
    constexpr int seg = rem < room ? rem : room;
    if (r.mine) st_segment<NPAIR, OFF, seg>(xk2, xv2, r.base + r.as * TP_SLOT_COLS + CIN);
    if constexpr (CIN + seg == TP_SLOT_COLS) {
      ring2_end(r);
      emit_field2<NPAIR, 0, OFF + seg>(xk2, xv2, r);
    }
  }
}

template <int NI8>
__global__ void __launch_bounds__(TP_THREADS, 1)
cin_fwd_tc_dual_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmXk, CinTcParams p) {
  constexpr int HpP = NI8 * 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_box_bytes = (uint32_t)p.H_pad * 128;                     // one 64-wide K chunk
  const uint32_t w_stage_bytes = 2 * w_box_bytes;
  uint8_t* sW = smem;                                                       // ns_w x 2 x [H_pad x 128 B]
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;
  const uint32_t xk_tile = (uint32_t)128 * HpP * 2;
  uint8_t* sX0 = sW + (size_t)p.ns_w * w_stage_bytes;                       // 2 buffers
  uint8_t* sXk = sX0 + 2 * (size_t)x0_tile;                                 // 2 buffers
  float* sBias = reinterpret_cast<float*>(sXk + 2 * (size_t)xk_tile);       // [H_pad]
  CinTpBars* bars = reinterpret_cast<CinTpBars*>(sBias + p.H_pad);

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TP_MAX_NS_W; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < TP_ASLOTS; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->x_full[i], 1); mbar_init(&bars->x_empty[i], 8); }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, 8);
    fence_barrier_init();
  }
  for (int h = threadIdx.x; h < p.H_pad; h += TP_THREADS) sBias[h] = h < p.H ? p.bias[h] : 0.f;
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int n_stages = p.n_wchunks;                                         // 128-wide K stages per tile
  auto tile_of = [&](int it) -> int64_t { return (int64_t)it * gridDim.x + blockIdx.x; };

  if (warp == 0) {
    // =============================== TMA: x tiles (one tile ahead) and the W' stage stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmXk);
      const int slice = p.H_pad / (int)csize;
      const int wr0 = (int)crank * slice;
      int xit = 0;
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
      uint32_t ws = 0, wphase = 1;
      bool first_pass = true;
      if (tile_of(0) < p.n_tiles) load_x(tile_of(0));
      for (int it = 0; it < p.n_iters; ++it) {
        if (it + 1 < p.n_iters && tile_of(it + 1) < p.n_tiles) load_x(tile_of(it + 1));
        for (int c = 0; c < n_stages; ++c) {
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
    // =============================== MMA issuer: per stage wait W' + A, eight MMAs, release both ===============================
    const uint32_t idesc = make_idesc_bf16(128, p.H_pad);
    const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
    const uint32_t stage_desc_step = w_stage_bytes >> 4;
    const uint32_t box_desc_step = w_box_bytes >> 4;
    uint32_t ws = 0, wphase = 0;
    uint64_t bdesc = bdesc0;
    uint32_t g = 0;
    int at = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      const bool active = tile_of(it) < p.n_tiles;
      if (active && at > 0) {
        mbar_wait(&bars->acc_empty, (at - 1) & 1);
        fence_after_sync();
      }
      for (int c = 0; c < n_stages; ++c) {
        mbar_wait(&bars->w_full[ws], wphase);
        const uint32_t as = g & (TP_ASLOTS - 1);
        if (active) mbar_wait(&bars->a_full[as], (g / TP_ASLOTS) & 1);
        fence_after_sync();
        if (elect_one()) {
          if (active) {
            const uint32_t a_addr = tmem_base + TP_A_COL0 + as * TP_SLOT_COLS;
            const uint64_t bdesc1 = bdesc + box_desc_step;
            umma_ts(tmem_base, a_addr, bdesc, idesc, c > 0 ? 1u : 0u);
            umma_ts(tmem_base, a_addr + 8, bdesc + 2, idesc, 1u);
            umma_ts(tmem_base, a_addr + 16, bdesc + 4, idesc, 1u);
            umma_ts(tmem_base, a_addr + 24, bdesc + 6, idesc, 1u);
            umma_ts(tmem_base, a_addr + 32, bdesc1, idesc, 1u);
            umma_ts(tmem_base, a_addr + 40, bdesc1 + 2, idesc, 1u);
            umma_ts(tmem_base, a_addr + 48, bdesc1 + 4, idesc, 1u);
            umma_ts(tmem_base, a_addr + 56, bdesc1 + 6, idesc, 1u);
            umma_commit(&bars->a_empty[as]);
          }
          if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
          else umma_commit(&bars->w_empty[ws]);
        }
        __syncwarp();
        if (++ws == (uint32_t)p.ns_w) { ws = 0; wphase ^= 1; bdesc = bdesc0; }
        else bdesc += stage_desc_step;
        if (active) ++g;
      }
      if (active) {
        if (elect_one()) umma_commit(&bars->acc_full);
        __syncwarp();
        ++at;
      }
    }
  } else {
    // =============================== row warps: Z producer (alternate slots), then epilogue (column half) ===============================
    const int q = warp & 3;                      // TMEM lane quarter
    const int par = (warp - 2) >> 2;             // producer slot parity / epilogue column half
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    ARing2 ring;
    ring.bars = bars;
    ring.base = tmem_base + lane_addr + TP_A_COL0;
    ring.g = 0;
    ring.parity = (uint32_t)par;
    ring.as = 0;
    ring.mine = false;
    ring.lane = lane;
    int at = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      const int64_t tile = tile_of(it);
      if (tile >= p.n_tiles) continue;
      const int buf = at & 1;
      mbar_wait(&bars->x_full[buf], (at >> 1) & 1);
      const uint4* xkrow = reinterpret_cast<const uint4*>(sXk + (size_t)buf * xk_tile + (size_t)rl * HpP * 2);
      __nv_bfloat162 xk2[HpP / 2];
#pragma unroll
      for (int v8 = 0; v8 < NI8; ++v8) {
        const uint4 t = xkrow[v8];
        xk2[v8 * 4 + 0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
        xk2[v8 * 4 + 1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        xk2[v8 * 4 + 2] = *reinterpret_cast<const __nv_bfloat162*>(&t.z);
        xk2[v8 * 4 + 3] = *reinterpret_cast<const __nv_bfloat162*>(&t.w);
      }
      const __nv_bfloat16* x0row = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)buf * x0_tile) + (size_t)rl * p.mP;
      int ph = 0;                                // 4-column granule inside the current slot where the next field starts
      for (int j = 0; j < p.m; ++j) {
        const __nv_bfloat16 xv = x0row[j];
        const __nv_bfloat162 xv2 = __halves2bfloat162(xv, xv);
        switch (ph) {
          case 0: emit_field2<HpP / 2, 0, 0>(xk2, xv2, ring); break;
          case 1: emit_field2<HpP / 2, 4, 0>(xk2, xv2, ring); break;
          case 2: emit_field2<HpP / 2, 8, 0>(xk2, xv2, ring); break;
          case 3: emit_field2<HpP / 2, 12, 0>(xk2, xv2, ring); break;
          case 4: emit_field2<HpP / 2, 16, 0>(xk2, xv2, ring); break;
          case 5: emit_field2<HpP / 2, 20, 0>(xk2, xv2, ring); break;
          case 6: emit_field2<HpP / 2, 24, 0>(xk2, xv2, ring); break;
          case 7: emit_field2<HpP / 2, 28, 0>(xk2, xv2, ring); break;
          case 8: emit_field2<HpP / 2, 32, 0>(xk2, xv2, ring); break;
          case 9: emit_field2<HpP / 2, 36, 0>(xk2, xv2, ring); break;
          case 10: emit_field2<HpP / 2, 40, 0>(xk2, xv2, ring); break;
          case 11: emit_field2<HpP / 2, 44, 0>(xk2, xv2, ring); break;
          case 12: emit_field2<HpP / 2, 48, 0>(xk2, xv2, ring); break;
          case 13: emit_field2<HpP / 2, 52, 0>(xk2, xv2, ring); break;
          case 14: emit_field2<HpP / 2, 56, 0>(xk2, xv2, ring); break;
          default: emit_field2<HpP / 2, 60, 0>(xk2, xv2, ring); break;
        }
        ph = (ph + NI8) & 15;
      }
      if (ph != 0) {                             // zero-fill the tail of the last K stage and publish it
        if (ring.mine) {
          const uint32_t zz[4] = {0u, 0u, 0u, 0u};
          for (; ph < 16; ++ph) tmem_st_x4(ring.base + ring.as * TP_SLOT_COLS + ph * 4, zz);
        }
        ring2_end(ring);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_empty[buf]);
      // ---- epilogue of this tile (this warp: 16-column chunks par*16, +32, ...)
      const int64_t row = tile * 128 + rl;
      const int64_t b = row / p.D;
      const int d = (int)(row - b * p.D);
      EpiRow r;
      r.valid = row < p.R;
      r.yt = (p.yt && r.valid) ? p.yt + row * p.Hs : nullptr;
      r.maps = (p.maps && r.valid) ? p.maps + (b * p.fm_total + p.col_off - p.hdb) * (int64_t)p.D + d : nullptr;
      r.pooled = (p.pooled && r.valid) ? p.pooled + b * p.fm_total + p.col_off - p.hdb : nullptr;
      r.spool = nullptr;
      r.act_floor = p.act == XDFM_ACT_RELU ? 0.f : __int_as_float(0xff800000);
      mbar_wait(&bars->acc_full, at & 1);
      fence_after_sync();
      const uint32_t acc = tmem_base + lane_addr;
      if (p.D == 8) epilogue_tile_act<8>(acc, par, lane, sBias, p, r);
      else if (p.D == 16) epilogue_tile_act<16>(acc, par, lane, sBias, p, r);
      else epilogue_tile_act<32>(acc, par, lane, sBias, p, r);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);
      ++at;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}
