// CIN forward, "tile pair" variant: every CTA contracts TWO 128-row tiles against each streamed W' chunk.
//
// Why: ncu / switch-off experiments of round 1 showed the single-tile kernels limited by the W' stream, not by the tensor core:
// every 128-row tile re-reads the whole W' (1.1 MB at cfg2) from L2, ~7 TB/s aggregate SM ingest at 148 CTAs, while the MMAs of
// a tile only need 18 k cycles.  With two accumulators per CTA ([128 x H_pad] each, H_pad <= 224 so that both plus an A ring fit
// the 512 TMEM columns) a chunk is fetched once per 256 rows: half the stream, eight MMAs per W' barrier round trip.
//
// TMEM: [0, H_pad) accumulator of tile 0, [H_pad, 2 H_pad) accumulator of tile 1, [2 H_pad, 512) ring of NS A slots of 32 columns
// (64 K-values), shared by the two tiles' producer streams in the order (chunk 0, tile 0), (chunk 0, tile 1), (chunk 1, tile 0)...
// Warps (10): 0 = TMA, 1 = MMA issuer + TMEM alloc, 2..5 = rows of tile 0, 6..9 = rows of tile 1 (warp & 3 = TMEM lane quarter).
// A row warp first generates its tile's Z chunks, then -- the TMEM being full, nothing can overlap anyway -- drains its tile's
// accumulator (bias / activation / bf16 store / split-half sum-pool), so producers and epilogue share the same 8 warps.
// Tile schedule: iteration `it` of CTA b owns tiles it*2G + b and it*2G + G + b (G = grid size); an iteration may hold 2, 1 or 0
// tiles (tail), every role derives the same count, and the W' ring advances identically in all cases (cluster lock-step).
// Included by cin_tc.cu (uses its Z-producer / epilogue helpers).  D <= 32 only (the D > 32 pooling needs a cross-warp pass).
#pragma once

#define TP_THREADS 320
#define TP_SLOT_COLS 32
#define TP_MAX_NS_W 6
#define TP_MAX_ASLOTS 8

struct __align__(8) CinTpBars {
  uint64_t w_full[TP_MAX_NS_W], w_empty[TP_MAX_NS_W];
  uint64_t a_full[TP_MAX_ASLOTS], a_empty[TP_MAX_ASLOTS];
  uint64_t x_full[2], x_empty[2];          // per tile slot (single-buffered: rows are copied to registers at tile start)
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct ARing2 {
  CinTpBars* bars;
  uint32_t base;      // TMEM address of ring column 0 for this warp's lane quarter
  uint32_t g;         // global slot sequence number of the chunk being written
  uint32_t ns;        // ring slots
  uint32_t stride;    // streams sharing the ring in this iteration (1 or 2)
  uint32_t as;        // g % ns
  int lane;
};

__device__ __forceinline__ void ring2_acquire(ARing2& r) {
  r.as = r.g % r.ns;
  const uint32_t n = r.g / r.ns;
  if (n > 0) {
    mbar_wait(&r.bars->a_empty[r.as], (n - 1) & 1);
    fence_after_sync();
  }
}
__device__ __forceinline__ void ring2_publish(ARing2& r) {
  tmem_wait_st();
  fence_before_sync();
  __syncwarp();
  if (r.lane == 0) mbar_arrive(&r.bars->a_full[r.as]);
  r.g += r.stride;
}

template <int NPAIR, int CIN, int OFF>
__device__ __forceinline__ void emit_field2(const __nv_bfloat162 (&xk2)[NPAIR], __nv_bfloat162 xv2, ARing2& r) {
  if constexpr (OFF < NPAIR) {
    if constexpr (CIN == 0) ring2_acquire(r);
    constexpr int room = TP_SLOT_COLS - CIN;
    constexpr int rem = NPAIR - OFF;
    constexpr int seg = rem < room ? rem : room;
    st_segment<NPAIR, OFF, seg>(xk2, xv2, r.base + r.as * TP_SLOT_COLS + CIN);
    if constexpr (CIN + seg == TP_SLOT_COLS) {
      ring2_publish(r);
      emit_field2<NPAIR, 0, OFF + seg>(xk2, xv2, r);
    }
  }
}

template <int NI8>
__global__ void __launch_bounds__(TP_THREADS, 1)
cin_fwd_tc_pair_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmXk, CinTcParams p) {
  constexpr int HpP = NI8 * 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_stage_bytes = (uint32_t)p.H_pad * 128;                   // one 64-wide K chunk
  uint8_t* sW = smem;                                                       // ns_w x [H_pad x 128 B]
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;
  const uint32_t xk_tile = (uint32_t)128 * HpP * 2;
  uint8_t* sX0 = sW + (size_t)p.ns_w * w_stage_bytes;                       // 2 tiles
  uint8_t* sXk = sX0 + 2 * (size_t)x0_tile;                                 // 2 tiles
  float* sBias = reinterpret_cast<float*>(sXk + 2 * (size_t)xk_tile);       // [H_pad]
  CinTpBars* bars = reinterpret_cast<CinTpBars*>(sBias + p.H_pad);
  const uint32_t ns_a = (512u - 2u * (uint32_t)p.H_pad) / TP_SLOT_COLS < TP_MAX_ASLOTS ? (512u - 2u * (uint32_t)p.H_pad) / TP_SLOT_COLS
                                                                                          : TP_MAX_ASLOTS;
  const uint32_t ring_col0 = 2u * (uint32_t)p.H_pad;
  const int n_chunks = p.n_wchunks;                                         // 64-wide chunks here

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TP_MAX_NS_W; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < TP_MAX_ASLOTS; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->x_full[i], 1);   mbar_init(&bars->x_empty[i], 4);
      mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  for (int h = threadIdx.x; h < p.H_pad; h += TP_THREADS) sBias[h] = h < p.H ? p.bias[h] : 0.f;
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  const int64_t G = gridDim.x;
  auto tile_of = [&](int it, int s) -> int64_t { return (int64_t)it * 2 * G + (int64_t)s * G + blockIdx.x; };
  auto ntiles_of = [&](int it) -> int { return tile_of(it, 1) < p.n_tiles ? 2 : (tile_of(it, 0) < p.n_tiles ? 1 : 0); };

  if (warp == 0) {
    // =============================== TMA: x rows of the iteration's tiles, then the W' chunk stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmXk);
      const int slice = p.H_pad / (int)csize;
      const int wr0 = (int)crank * slice;
      uint32_t ws = 0, wphase = 1;
      bool first_pass = true;
      int xuse[2] = {0, 0};
      for (int it = 0; it < p.n_iters; ++it) {
        const int nt = ntiles_of(it);
        for (int s = 0; s < nt; ++s) {
          if (xuse[s] > 0) mbar_wait(&bars->x_empty[s], (xuse[s] - 1) & 1);
          const int64_t r0 = tile_of(it, s) * 128;
          const uint32_t nrows = (uint32_t)min((int64_t)128, p.R - r0);
          mbar_arrive_expect_tx(&bars->x_full[s], nrows * (uint32_t)(p.mP * 2) + xk_tile);
          bulk_load_1d(sX0 + (size_t)s * x0_tile, p.x0t + r0 * p.mP, nrows * (uint32_t)(p.mP * 2), &bars->x_full[s]);
          tma_load_2d(sXk + (size_t)s * xk_tile, &tmXk, 0, (int)r0, &bars->x_full[s]);   // OOB rows are zero-filled
          ++xuse[s];
        }
        for (int c = 0; c < n_chunks; ++c) {
          if (!first_pass) mbar_wait(&bars->w_empty[ws], wphase);
          mbar_arrive_expect_tx(&bars->w_full[ws], w_stage_bytes);
          uint8_t* dst = sW + (size_t)ws * w_stage_bytes + (size_t)wr0 * 128;
          if (csize > 1) tma_load_2d_mcast(dst, &tmW, c * 64, wr0, &bars->w_full[ws], cmask);
          else tma_load_2d(dst, &tmW, c * 64, wr0, &bars->w_full[ws]);
          if (++ws == (uint32_t)p.ns_w) { ws = 0; wphase ^= 1; first_pass = false; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = make_idesc_bf16(128, p.H_pad);
    const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
    const uint32_t stage_desc_step = w_stage_bytes >> 4;
    uint32_t ws = 0, wphase = 0;
    uint64_t bdesc = bdesc0;
    uint32_t g = 0;                                             // A ring sequence number
    int use[2] = {0, 0};                                        // accumulator uses so far
    for (int it = 0; it < p.n_iters; ++it) {
      const int nt = ntiles_of(it);
      for (int s = 0; s < nt; ++s) {
        if (use[s] > 0) {
          mbar_wait(&bars->acc_empty[s], (use[s] - 1) & 1);
          fence_after_sync();
        }
      }
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(&bars->w_full[ws], wphase);
        for (int s = 0; s < nt; ++s) {
          const uint32_t as = g % ns_a, n = g / ns_a;
          mbar_wait(&bars->a_full[as], n & 1);
          fence_after_sync();
          if (elect_one()) {
            const uint32_t d_addr = tmem_base + (uint32_t)s * (uint32_t)p.H_pad;
            const uint32_t a_addr = tmem_base + ring_col0 + as * TP_SLOT_COLS;
            umma_ts(d_addr, a_addr, bdesc, idesc, c > 0 ? 1u : 0u);
            umma_ts(d_addr, a_addr + 8, bdesc + 2, idesc, 1u);
            umma_ts(d_addr, a_addr + 16, bdesc + 4, idesc, 1u);
            umma_ts(d_addr, a_addr + 24, bdesc + 6, idesc, 1u);
            umma_commit(&bars->a_empty[as]);
          }
          __syncwarp();
          ++g;
        }
        if (nt == 0) fence_after_sync();
        if (elect_one()) {
          if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
          else umma_commit(&bars->w_empty[ws]);
        }
        __syncwarp();
        if (++ws == (uint32_t)p.ns_w) { ws = 0; wphase ^= 1; bdesc = bdesc0; }
        else bdesc += stage_desc_step;
      }
      for (int s = 0; s < nt; ++s) {
        if (elect_one()) umma_commit(&bars->acc_full[s]);
        __syncwarp();
        ++use[s];
      }
    }
  } else {
    // =============================== row warps: Z producer, then epilogue, of one tile slot ===============================
    const int s = (warp - 2) >> 2;               // tile slot 0 / 1
    const int q = warp & 3;                      // TMEM lane quarter
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    ARing2 ring;
    ring.bars = bars;
    ring.base = tmem_base + lane_addr + ring_col0;
    ring.ns = ns_a;
    ring.lane = lane;
    uint32_t g_base = 0;                         // ring sequence number of (chunk 0, tile 0) of the current iteration
    int use = 0;
    for (int it = 0; it < p.n_iters; ++it) {
      const int nt = ntiles_of(it);
      if (s < nt) {
        const int64_t tile = tile_of(it, s);
        // ---- operand rows -> registers (the shared-memory tile is handed back to the TMA warp immediately)
        mbar_wait(&bars->x_full[s], use & 1);
        const uint4* xkrow = reinterpret_cast<const uint4*>(sXk + (size_t)s * xk_tile + (size_t)rl * HpP * 2);
        __nv_bfloat162 xk2[HpP / 2];
#pragma unroll
        for (int v8 = 0; v8 < NI8; ++v8) {
          const uint4 t = xkrow[v8];
          xk2[v8 * 4 + 0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
          xk2[v8 * 4 + 1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
          xk2[v8 * 4 + 2] = *reinterpret_cast<const __nv_bfloat162*>(&t.z);
          xk2[v8 * 4 + 3] = *reinterpret_cast<const __nv_bfloat162*>(&t.w);
        }
        const __nv_bfloat16* x0row = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)s * x0_tile) + (size_t)rl * p.mP;
        ring.g = g_base + (uint32_t)s;
        ring.stride = (uint32_t)nt;
        int ph = 0;
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
            default: emit_field2<HpP / 2, 28, 0>(xk2, xv2, ring); break;
          }
          ph = (ph + NI8) & 7;
        }
        if (ph != 0) {                             // zero-fill the tail of the last K chunk and publish it
          const uint32_t zz[4] = {0u, 0u, 0u, 0u};
          for (; ph < 8; ++ph) tmem_st_x4(ring.base + ring.as * TP_SLOT_COLS + ph * 4, zz);
          ring2_publish(ring);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->x_empty[s]);     // x0row is not read after this point
        // ---- epilogue of this tile
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
        mbar_wait(&bars->acc_full[s], use & 1);
        fence_after_sync();
        const uint32_t acc = tmem_base + lane_addr + (uint32_t)s * (uint32_t)p.H_pad;
        if (p.D == 8) { epilogue_tile_act<8>(acc, 0, lane, sBias, p, r); epilogue_tile_act<8>(acc, 1, lane, sBias, p, r); }
        else if (p.D == 16) { epilogue_tile_act<16>(acc, 0, lane, sBias, p, r); epilogue_tile_act<16>(acc, 1, lane, sBias, p, r); }
        else { epilogue_tile_act<32>(acc, 0, lane, sBias, p, r); epilogue_tile_act<32>(acc, 1, lane, sBias, p, r); }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[s]);
        ++use;
      }
      g_base += (uint32_t)(n_chunks * nt);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}
