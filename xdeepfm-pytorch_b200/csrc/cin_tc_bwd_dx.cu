// CIN layer backward w.r.t. the activations on the tensor cores.
//
// Replaces the dZ / dX part of torch's autograd of deepctr/layers/interaction.py:218-224 (convolution_backward input grad +
// einsum backward).  With dY = act'(y) * upstream (bf16, row layout [R, Hs], R = B*D rows r = (sample, d)):
//
//     dZ[r, (j,i)] = sum_h dY[r,h] * W[h, i*m + j]                 (implicit GEMM, never materialised)
//     dXk[r, i]    = sum_j dZ[r,(j,i)] * X0[r, j]                   (layer input gradient, fp32 row layout [R, HpQ])
//     dX0[r, j]   += sum_i dZ[r,(j,i)] * Xk[r, i]                   (accumulated over layers, fp32 row layout [R, mP])
//
// One accumulator tile = 128 rows.  A = dY tile, written to TMEM once per tile by the row warps (tcgen05.st, TS-mode MMA).
// For every X^0 field j one MMA group computes dZ_j[128 x HpQ] = dY[128 x H_pad] . W''_j[HpQ x H_pad]^T into one of two TMEM
// accumulators (HpQ = Hp rounded up to 16); while the tensor core works on field j+1 the row warps drain field j:
// each thread owns one row, multiplies the chunk by X0[r,j] into its register-resident dXk row and dots it with its
// register-resident Xk row for dX0[r,j].  Two warps share a TMEM lane quarter and split the channel range.
// W'' = weights permuted to [m*HpQ rows (j-major), H_pad cols] bf16 (K-major for this GEMM), streamed by TMA per field and
// multicast across the CTAs of a cluster exactly like the forward weight stream.
//
// TMEM columns: [0,128) / [128,256) dY of the current / next tile (A operand, double buffered), [256,384) / [384,512) the two dZ
// accumulators.  The row warps stage the NEXT tile's dY one 16-byte granule per field while they drain the current tile (global
// load issued at field g, tcgen05.st at field g+1), so the tensor core rolls from the last field of a tile straight into the
// first field of the next one; shapes with fewer fields than granules stage at the tile start instead.
// Warps: 0 = TMA, 1 = MMA + TMEM alloc, 2..9 = row warps (quarter = warp & 3, channel half = (warp - 2) >> 2).
#include "tc_common.cuh"
#include "../../include/xdfm.h"

using namespace tc;

#define DX_THREADS 320
#define DX_ACC_COL0 256
#define DX_ACC_COLS 128
#define DX_MAX_NS 8

struct CinDxParams {
  const __nv_bfloat16* dyt;   // [R, Hs]
  const __nv_bfloat16* x0t;   // [R, mP]
  const __nv_bfloat16* xkt;   // rows with pitch xk_pitch, first Hp channels used
  float* dxk;                 // [R, HpQ] fp32 (overwritten)
  float* dx0;                 // [R, mP]  fp32 (accumulated)
  int64_t R, xk_pitch;
  int m, mP, Hp, HpQ, H, H_pad, Hs;
  int64_t n_tiles;
  int n_iters;
  int n_hchunks;              // 64-wide chunks of the reduction dim h per field = ceil(H_pad / 64)
  int debug;                  // diagnostic bit mask (0 in production): 1 skip contraction FMAs, 2 skip TMEM loads, 4 skip MMAs
  int ns;                     // W'' ring depth (one slot = one FIELD on one barrier)
  // single-tile kernel: a slot holds the field's n_full 64-wide h-chunks ([HpQ rows x 128 B], SWIZZLE_128B) followed by tail_ks
  // 16-wide chunks ([HpQ rows x 32 B], SWIZZLE_32B) -- H_pad = 64 n_full + 16 tail_ks, nothing zero-padded is streamed
  int n_full, tail_ks;        // the ring keeps the full chunks of all slots first (1024-byte aligned), then the tails (256-byte aligned)
  int na_shift;               // log2 of the number of dZ accumulators in flight: 2 x 128 columns, or 4 x 64 when HpQ <= 64
};

struct __align__(8) CinDxBars {
  uint64_t w_full[DX_MAX_NS], w_empty[DX_MAX_NS];
  uint64_t a_full[2], a_empty[2];    // dY tiles in TMEM (count 4 * NG / 1)
  uint64_t acc_full[4], acc_empty[4];
  uint64_t x_full[2], x_empty[2];
  uint32_t tmem_base;
};

// NQ = HpQ / 16; NG = row warps per TMEM lane quarter (2 or 4): each drains HpQ / NG channels of every dZ_j.  ncu (round 1) shows the
// row warps, not the tensor core, pacing this kernel at low IPC (TMEM-load and FMA latencies, spills at 168 registers); with NG = 4
// there are twice as many resident warps to hide those latencies and each needs half the registers.
template <int NQ, int NG>
__global__ void __launch_bounds__((2 + 4 * NG) * 32, 1)
cin_bwd_dx_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWt, CinDxParams p) {
  constexpr int HpQ = NQ * 16;
  constexpr int HALF = HpQ / NG;                   // channels per row warp (multiple of 4)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_box_bytes = (uint32_t)HpQ * 128;                          // one 64-wide h-chunk of one field
  const uint32_t w_tail_bytes = (uint32_t)HpQ * 32;                          // one 16-wide h-chunk of the tail
  const uint32_t full_stride = w_box_bytes * (uint32_t)p.n_full;             // a slot's full chunks
  const uint32_t tail_stride = w_tail_bytes * (uint32_t)p.tail_ks;           // a slot's tail chunks
  const uint32_t w_slot_bytes = full_stride + tail_stride;                    // bytes TMA delivers per field
  uint8_t* sW = smem;                                                         // ns x full chunks
  uint8_t* sWt = sW + (size_t)p.ns * full_stride;                             // ns x tail chunks
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;
  uint8_t* sX0 = sWt + (size_t)p.ns * tail_stride;                            // 2 x [128][mP] bf16
  const int dpitch = p.m | 1;                                                 // odd pitch: a warp's 32 rows hit 32 different banks
  float* sDx0 = reinterpret_cast<float*>(sX0 + 2 * (size_t)x0_tile);          // [NG groups][128][dpitch] fp32 dX0 partials
  CinDxBars* bars = reinterpret_cast<CinDxBars*>(sDx0 + ((NG * 128 * dpitch + 1) & ~1));

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < DX_MAX_NS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4 * NG); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 4 * NG); mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->x_full[i], 1);   mbar_init(&bars->x_empty[i], 4 * NG);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  auto tile_of = [&](int it) -> int64_t { return (int64_t)it * gridDim.x + blockIdx.x; };

  if (warp == 0) {
    // =============================== TMA: x0 tiles (one ahead) + W'' stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      prefetch_tmap(&tmWt);
      const int slice = HpQ / (int)csize;            // rows of a W'' slot loaded (and multicast) by this CTA
      const int wr0 = (int)crank * slice;
      int xit = 0;
      auto load_x = [&](int64_t tile) {
        const int buf = xit & 1;
        if (xit >= 2) mbar_wait(&bars->x_empty[buf], ((xit >> 1) - 1) & 1);
        const int64_t r0 = tile * 128;
        const uint32_t nrows = (uint32_t)min((int64_t)128, p.R - r0);
        mbar_arrive_expect_tx(&bars->x_full[buf], nrows * (uint32_t)(p.mP * 2));
        bulk_load_1d(sX0 + (size_t)buf * x0_tile, p.x0t + r0 * p.mP, nrows * (uint32_t)(p.mP * 2), &bars->x_full[buf]);
        ++xit;
      };
      uint32_t ws = 0, wphase = 1;
      bool first_pass = true;
      if (tile_of(0) < p.n_tiles) load_x(tile_of(0));
      for (int it = 0; it < p.n_iters; ++it) {
        if (it + 1 < p.n_iters && tile_of(it + 1) < p.n_tiles) load_x(tile_of(it + 1));
        for (int j = 0; j < p.m; ++j) {
          if (!first_pass) mbar_wait(&bars->w_empty[ws], wphase);
          if (p.debug & 8) {                      // experiment: no weight stream at all (only the barrier hand-offs remain)
            mbar_arrive(&bars->w_full[ws]);
            if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; first_pass = false; }
            continue;
          }
          mbar_arrive_expect_tx(&bars->w_full[ws], w_slot_bytes);
          for (int c = 0; c < p.n_full; ++c) {
            uint8_t* dst = sW + (size_t)ws * full_stride + (size_t)c * w_box_bytes + (size_t)wr0 * 128;
            if (csize > 1) tma_load_2d_mcast(dst, &tmW, c * 64, j * HpQ + wr0, &bars->w_full[ws], cmask);
            else tma_load_2d(dst, &tmW, c * 64, j * HpQ + wr0, &bars->w_full[ws]);
          }
          for (int t = 0; t < p.tail_ks; ++t) {
            uint8_t* dst = sWt + (size_t)ws * tail_stride + (size_t)t * w_tail_bytes + (size_t)wr0 * 32;
            if (csize > 1) tma_load_2d_mcast(dst, &tmWt, p.n_full * 64 + t * 16, j * HpQ + wr0, &bars->w_full[ws], cmask);
            else tma_load_2d(dst, &tmWt, p.n_full * 64 + t * 16, j * HpQ + wr0, &bars->w_full[ws]);
          }
          if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; first_pass = false; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform loop, elected lane issues) ===============================
    const uint32_t idesc = make_idesc_bf16(128, HpQ);
    const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
    const uint64_t tdesc0 = make_desc_k_sw32(smem_u32(sWt));
    const uint32_t slot_desc_step = full_stride >> 4;
    const uint32_t tslot_desc_step = tail_stride >> 4;
    const uint32_t box_desc_step = w_box_bytes >> 4;
    const uint32_t tail_desc_step = w_tail_bytes >> 4;
    uint32_t ws = 0, wphase = 0;
    uint64_t bdesc = bdesc0, tdesc = tdesc0;
    uint32_t jc = 0;            // fields processed so far (accumulator buffer = jc & na_mask)
    int at = 0;
    const int ksteps = p.n_full * 4;
    const uint32_t na_mask = (1u << p.na_shift) - 1, acc_stride = 256u >> p.na_shift;
    for (int it = 0; it < p.n_iters; ++it) {
      const bool active = tile_of(it) < p.n_tiles;
      const uint32_t abuf = (uint32_t)(at & 1);
      const uint32_t a_addr0 = tmem_base + abuf * 128;
      if (active) {
        mbar_wait(&bars->a_full[abuf], (at >> 1) & 1);
        fence_after_sync();
      }
      for (int j = 0; j < p.m; ++j) {
        const uint32_t ab = jc & na_mask;
        if (active && jc > na_mask) {
          mbar_wait(&bars->acc_empty[ab], ((jc >> p.na_shift) - 1) & 1);
          fence_after_sync();
        }
        mbar_wait(&bars->w_full[ws], wphase);
        fence_after_sync();
        if (elect_one()) {
          if (active && !(p.debug & 4)) {
            const uint32_t d_addr = tmem_base + DX_ACC_COL0 + ab * acc_stride;
            uint64_t bd = bdesc;
            for (int ks = 0; ks < ksteps; ks += 4, bd += box_desc_step) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                if (ks + k4 < ksteps) umma_ts(d_addr, a_addr0 + (uint32_t)(ks + k4) * 8, bd + (uint64_t)(k4 * 2), idesc, (ks + k4) > 0 ? 1u : 0u);
              }
            }
            for (int t = 0; t < p.tail_ks; ++t)
              umma_ts(d_addr, a_addr0 + (uint32_t)(ksteps + t) * 8, tdesc + (uint64_t)t * tail_desc_step, idesc, (ksteps + t) > 0 ? 1u : 0u);
          }
          if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
          else umma_commit(&bars->w_empty[ws]);
          if (active) umma_commit(&bars->acc_full[ab]);
        }
        __syncwarp();
        if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; bdesc = bdesc0; tdesc = tdesc0; }
        else { bdesc += slot_desc_step; tdesc += tslot_desc_step; }
        if (active) ++jc;
      }
      if (active) {
        if (elect_one()) umma_commit(&bars->a_empty[abuf]);   // all MMAs reading this tile's dY have been issued and will complete
        __syncwarp();
        ++at;
      }
    }
  } else {
    // =============================== row warps ===============================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;              // channel group: channels [half * HALF, (half + 1) * HALF)
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t jc = 0;
    int at = 0;
    const uint32_t na_mask = (1u << p.na_shift) - 1, acc_stride = 256u >> p.na_shift;
    bool staged = false;        // this tile's dY already sits in TMEM (staged while the previous tile was drained)
    for (int it = 0; it < p.n_iters; ++it) {
      const int64_t tile = tile_of(it);
      if (tile >= p.n_tiles) continue;
      const int64_t row = tile * 128 + rl;
      const bool valid = row < p.R;
      // ---- this tile's dY rows in TMEM (A operand, buffer at & 1): this warp owns 32-bit columns [c_beg, c_end) of its lane quarter.
      // Staged during the previous tile when the shape allows it (one granule per field), else here.
      const int ncol = p.H_pad / 2;                  // 32-bit columns of the A tile
      const int per = ((ncol + NG - 1) / NG + 3) & ~3;
      const int c_beg = half * per;
      const int c_end = min(ncol, c_beg + per);
      const int ngr = c_end > c_beg ? (c_end - c_beg + 3) / 4 : 0;     // this warp's 16-byte granules
      const uint32_t abuf = (uint32_t)(at & 1);
      if (!staged) {
        if (at >= 2) {
          mbar_wait(&bars->a_empty[abuf], ((at >> 1) - 1) & 1);
          fence_after_sync();
        }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.dyt + row * p.Hs);
        // all global loads of the row half are issued before the first TMEM store (ncu, round 1: 12.6 % of the kernel's samples sat
        // on the STTM of a load -> store loop that paid one global-load latency per 16 bytes); at most 17 granules (H_pad <= 256)
        constexpr int GB = 32 / NG + 1;
        uint4 gbuf[GB];
#pragma unroll
        for (int gi = 0; gi < GB; ++gi) {
          const int c = c_beg + gi * 4;
          gbuf[gi] = make_uint4(0u, 0u, 0u, 0u);
          if (c < c_end && valid && c * 2 < p.Hs) gbuf[gi] = *reinterpret_cast<const uint4*>(src + c);   // Hs multiple of 8: whole granules
        }
#pragma unroll
        for (int gi = 0; gi < GB; ++gi) {
          const int c = c_beg + gi * 4;
          if (c < c_end) {
            const uint32_t v[4] = {gbuf[gi].x, gbuf[gi].y, gbuf[gi].z, gbuf[gi].w};
            tmem_st_x4(tmem_base + lane_addr + abuf * 128 + c, v);
          }
        }
        tmem_wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_full[abuf]);
      }
      // next tile of this CTA: its dY is staged into the other A buffer while this tile's fields are drained
      const int64_t ntile = (it + 1 < p.n_iters) ? tile_of(it + 1) : p.n_tiles;
      const bool pipe = ntile < p.n_tiles && (per / 4) < p.m;          // granule g: load at field g, store at field g + 1 <= m - 1
      const uint32_t nbuf = abuf ^ 1u;
      const int64_t nrow = ntile * 128 + rl;
      const bool nvalid = pipe && nrow < p.R;
      const uint32_t* nsrc = reinterpret_cast<const uint32_t*>(p.dyt + (nvalid ? nrow : 0) * p.Hs);
      uint4 pg = make_uint4(0u, 0u, 0u, 0u);
      if (pipe && at >= 1) {                         // the other buffer was read by the previous tile's MMAs: long complete
        mbar_wait(&bars->a_empty[nbuf], (((at + 1) >> 1) - 1) & 1);
        fence_after_sync();
      }
      // ---- this row's operands
      const int buf = at & 1;
      mbar_wait(&bars->x_full[buf], (at >> 1) & 1);
      const __nv_bfloat16* x0row = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)buf * x0_tile) + (size_t)rl * p.mP;
      __nv_bfloat162 xk2[HALF / 2];
      {
        const __nv_bfloat16* xr = p.xkt + row * p.xk_pitch + half * HALF;
        if constexpr (HALF % 8 == 0) {
#pragma unroll
          for (int v8 = 0; v8 < HALF / 8; ++v8) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (valid && half * HALF + v8 * 8 < p.xk_pitch) t = *reinterpret_cast<const uint4*>(xr + v8 * 8);   // xk_pitch multiple of 8
            xk2[v8 * 4 + 0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
            xk2[v8 * 4 + 1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
            xk2[v8 * 4 + 2] = *reinterpret_cast<const __nv_bfloat162*>(&t.z);
            xk2[v8 * 4 + 3] = *reinterpret_cast<const __nv_bfloat162*>(&t.w);
          }
        } else {
#pragma unroll
          for (int v4 = 0; v4 < HALF / 4; ++v4) {
            uint2 t = make_uint2(0u, 0u);
            if (valid && half * HALF + v4 * 4 < p.xk_pitch) t = *reinterpret_cast<const uint2*>(xr + v4 * 4);
            xk2[v4 * 2 + 0] = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
            xk2[v4 * 2 + 1] = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
          }
        }
      }
      float dxk[HALF];
#pragma unroll
      for (int i = 0; i < HALF; ++i) dxk[i] = 0.f;
      float x0n = __bfloat162float(x0row[0]);
      for (int j = 0; j < p.m; ++j, ++jc) {
        const uint32_t ab = jc & na_mask;
        const float x0v = x0n;
        if (j + 1 < p.m) x0n = __bfloat162float(x0row[j + 1]);       // next field's scale: its shared-memory latency hides behind this field
        if (pipe && j <= ngr) {
          if (j >= 1) {                               // granule j - 1 was loaded one field ago
            const uint32_t v4[4] = {pg.x, pg.y, pg.z, pg.w};
            tmem_st_x4(tmem_base + lane_addr + nbuf * 128 + c_beg + (j - 1) * 4, v4);
          }
          if (j == ngr) {                             // this warp's share of the next A tile is complete
            tmem_wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->a_full[nbuf]);
          }
        }
        mbar_wait(&bars->acc_full[ab], (jc >> p.na_shift) & 1);
        fence_after_sync();
        const uint32_t acc = tmem_base + lane_addr + DX_ACC_COL0 + ab * acc_stride + half * HALF;
        float dot = 0.f;
        // this warp's dZ columns of the field in NB batches of TMEM loads (one wait each): one batch when the columns fit the register
        // budget next to the dXk accumulators, two otherwise (round 1: at HALF = 56 a single 56-register batch spilled the X^{k-1}
        // row into local memory inside this loop); the accumulator goes back to the tensor core after the last batch's loads
        constexpr int NB = HALF > 32 ? 2 : 1;
        constexpr int BS = HALF / NB;                 // multiple of 4
        static_assert(BS * NB == HALF && BS % 4 == 0, "dZ batch split");
        float d4[4] = {0.f, 0.f, 0.f, 0.f};          // four independent chains (the sum order is still a fixed function of HALF)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          uint32_t v[BS];
          if (p.debug & 2) {
#pragma unroll
            for (int i = 0; i < BS; ++i) v[i] = 0u;
          } else {
#pragma unroll
            for (int c0 = 0; c0 + 8 <= BS; c0 += 8) {
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                             "=r"(v[c0 + 6]), "=r"(v[c0 + 7])
                           : "r"(acc + nb * BS + c0)
                           : "memory");
            }
            if constexpr (BS % 8 != 0) {
              constexpr int c0 = BS - 4;
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3])
                           : "r"(acc + nb * BS + c0)
                           : "memory");
            }
          }
          tmem_wait_ld();
          if (nb == NB - 1) {
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[ab]);
            // the next tile's dY granule is requested only now: the arrive above has release semantics and waits for every
            // outstanding memory operation of the thread (ncu, round 1: 8 % of the kernel's samples were that MEMBAR with the
            // load in flight)
            if (pipe && j < ngr) {
              const int c = c_beg + j * 4;
              pg = make_uint4(0u, 0u, 0u, 0u);
              if (nvalid && c * 2 < p.Hs) pg = *reinterpret_cast<const uint4*>(nsrc + c);
            }
          }
          if (!(p.debug & 1)) {
#pragma unroll
            for (int i = 0; i < BS; i += 2) {
              const int ci = nb * BS + i;             // channel index inside this warp's half
              // bf16 pair -> two fp32 with volatile asm: the row is loop-invariant over the fields, and left to itself the compiler
              // hoists the conversions out of the field loop (56 more live registers -> the row spills to local memory)
              const uint32_t xp = *reinterpret_cast<const uint32_t*>(&xk2[ci / 2]);
              uint32_t xlo, xhi;
              asm volatile("shl.b32 %0, %1, 16;" : "=r"(xlo) : "r"(xp));
              asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(xhi) : "r"(xp));
              const float z0 = __uint_as_float(v[i]), z1 = __uint_as_float(v[i + 1]);
              dxk[ci] = fmaf(z0, x0v, dxk[ci]);
              dxk[ci + 1] = fmaf(z1, x0v, dxk[ci + 1]);
              d4[(ci / 2) & 3] = fmaf(z0, __uint_as_float(xlo), d4[(ci / 2) & 3]);
              d4[(ci / 2 + 2) & 3] = fmaf(z1, __uint_as_float(xhi), d4[(ci / 2 + 2) & 3]);
            }
          }
        }
        dot = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        // dX0[r, j] partial of this warp's channel group: parked in shared memory (plane = group), combined in group order at tile end
        sDx0[(half * 128 + rl) * dpitch + j] = dot;
      }
      staged = pipe;
      // ---- tile outputs
      if (valid && !(p.debug & 16)) {
        float* o = p.dxk + row * p.HpQ + half * HALF;
#pragma unroll
        for (int i = 0; i < HALF; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(dxk[i], dxk[i + 1], dxk[i + 2], dxk[i + 3]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NG) : "memory");     // all groups' dX0 partials are in shared memory
      if (valid && !(p.debug & 16)) {
        // dx0 row += the groups' partials (summed in group order): 128-bit accesses, all loads of a pass in flight before the first
        // add (the scalar load -> add -> store chain this replaces cost one L2 round trip per field at every tile end); the
        // channel groups split the row's float4s
        float4* g4 = reinterpret_cast<float4*>(p.dx0 + row * p.mP);
        const int nv4 = p.mP / 4;
        for (int base = half; base < nv4; base += NG * 4) {
          float4 buf[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = base + u * NG;
            if (i < nv4) buf[u] = g4[i];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = base + u * NG;
            if (i < nv4) {
              float a[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int j = i * 4 + t;
                if (j < p.m) {
                  float sacc = sDx0[rl * dpitch + j];
#pragma unroll
                  for (int gq = 1; gq < NG; ++gq) sacc += sDx0[(gq * 128 + rl) * dpitch + j];
                  a[t] += sacc;
                }
              }
              g4[i] = make_float4(a[0], a[1], a[2], a[3]);
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NG) : "memory");     // partial planes may be overwritten by the next tile
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->x_empty[buf]);
      ++at;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Tile-pair variant: every CTA contracts TWO 128-row tiles against each streamed W''_j.
//
// Switch-off experiments (profiles/r01_cin_findings.md) showed two thirds of the single-tile kernel's time to be its skeleton -- the
// W'' stream (1.5 GB per launch, ~7 TB/s of SM ingest) -- not MMAs, TMEM loads or FMAs.  Here a field's weights are fetched once per
// 256 rows: half the stream.  The two tiles' accumulators double-buffer each other (while the row warps drain tile 0's dZ_j the
// tensor core computes tile 1's), so TMEM holds A0 [0,128), A1 [128,256), acc0 [256,384), acc1 [384,512).  The same 8 row warps
// serve both tiles; to stay inside the register file their X^{k-1} row halves live in a thread-private shared-memory area
// ([tile][half][pair][row]: conflict-free) instead of registers, and each dZ chunk is drained in batches of 32 columns.  The two
// channel halves' dX0 partials meet in shared memory by atomicAdd onto a zeroed plane (two addends: order-independent).
// Tile schedule: iteration `it` of CTA b owns tiles it*2G + b and it*2G + G + b; 2, 1 or 0 of them exist.
// ------------------------------------------------------------------------------------------------
struct __align__(8) CinDxPairBars {
  uint64_t w_full[DX_MAX_NS], w_empty[DX_MAX_NS];
  uint64_t a_full[2], a_empty[2];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t x_full[2], x_empty[2];
  uint32_t tmem_base;
};

template <int NQ>
__global__ void __launch_bounds__(DX_THREADS, 1) cin_bwd_dx_tc_pair_kernel(const __grid_constant__ CUtensorMap tmW, CinDxParams p) {
  constexpr int HpQ = NQ * 16;
  constexpr int HALF = HpQ / 2;                    // channels per row warp (multiple of 8)
  constexpr int NPAIR = HALF / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_box_bytes = (uint32_t)HpQ * 128;
  const uint32_t w_slot_bytes = w_box_bytes * (uint32_t)p.n_hchunks;
  uint8_t* sW = smem;                                                         // ns x n_hchunks x [HpQ x 128 B]
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;
  uint8_t* sX0 = sW + (size_t)p.ns * w_slot_bytes;                            // 2 tiles x [128][mP] bf16
  uint32_t* sXk = reinterpret_cast<uint32_t*>(sX0 + 2 * (size_t)x0_tile);     // [2 tiles][2 halves][NPAIR][128] bf16x2 (thread-private)
  float* sDx0 = reinterpret_cast<float*>(sXk + 2 * 2 * NPAIR * 128);          // [2 tiles][128][mP] fp32, zero between tiles
  CinDxPairBars* bars = reinterpret_cast<CinDxPairBars*>(sDx0 + 2 * 128 * p.mP);

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < DX_MAX_NS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 8);   mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 8);
      mbar_init(&bars->x_full[i], 1);   mbar_init(&bars->x_empty[i], 8);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * 128 * p.mP; i += DX_THREADS) sDx0[i] = 0.f;
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  const int64_t G = gridDim.x;
  auto tile_of = [&](int it, int t) -> int64_t { return (int64_t)it * 2 * G + (int64_t)t * G + blockIdx.x; };
  auto ntiles_of = [&](int it) -> int { return tile_of(it, 1) < p.n_tiles ? 2 : (tile_of(it, 0) < p.n_tiles ? 1 : 0); };

  if (warp == 0) {
    // =============================== TMA: x0 rows of the iteration's tiles + the W'' stream ===============================
    if (lane == 0) {
      prefetch_tmap(&tmW);
      const int slice = HpQ / (int)csize;
      const int wr0 = (int)crank * slice;
      uint32_t ws = 0, wphase = 1;
      bool first_pass = true;
      int xuse[2] = {0, 0};
      for (int it = 0; it < p.n_iters; ++it) {
        const int nt = ntiles_of(it);
        for (int t = 0; t < nt; ++t) {
          if (xuse[t] > 0) mbar_wait(&bars->x_empty[t], (xuse[t] - 1) & 1);
          const int64_t r0 = tile_of(it, t) * 128;
          const uint32_t nrows = (uint32_t)min((int64_t)128, p.R - r0);
          mbar_arrive_expect_tx(&bars->x_full[t], nrows * (uint32_t)(p.mP * 2));
          bulk_load_1d(sX0 + (size_t)t * x0_tile, p.x0t + r0 * p.mP, nrows * (uint32_t)(p.mP * 2), &bars->x_full[t]);
          ++xuse[t];
        }
        for (int j = 0; j < p.m; ++j) {
          if (!first_pass) mbar_wait(&bars->w_empty[ws], wphase);
          mbar_arrive_expect_tx(&bars->w_full[ws], w_slot_bytes);
          for (int c = 0; c < p.n_hchunks; ++c) {
            uint8_t* dst = sW + (size_t)ws * w_slot_bytes + (size_t)c * w_box_bytes + (size_t)wr0 * 128;
            if (csize > 1) tma_load_2d_mcast(dst, &tmW, c * 64, j * HpQ + wr0, &bars->w_full[ws], cmask);
            else tma_load_2d(dst, &tmW, c * 64, j * HpQ + wr0, &bars->w_full[ws]);
          }
          if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; first_pass = false; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = make_idesc_bf16(128, HpQ);
    const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
    const uint32_t slot_desc_step = w_slot_bytes >> 4;
    const uint32_t box_desc_step = w_box_bytes >> 4;
    uint32_t ws = 0, wphase = 0;
    uint64_t bdesc = bdesc0;
    uint32_t fuse[2] = {0, 0};      // fields issued so far per tile slot (accumulator phase)
    int ause[2] = {0, 0};           // tiles staged so far per tile slot
    const int ksteps = p.H_pad / 16;
    for (int it = 0; it < p.n_iters; ++it) {
      const int nt = ntiles_of(it);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(&bars->a_full[t], ause[t] & 1);
        fence_after_sync();
      }
      for (int j = 0; j < p.m; ++j) {
        mbar_wait(&bars->w_full[ws], wphase);
        fence_after_sync();
        for (int t = 0; t < nt; ++t) {
          if (fuse[t] > 0) {
            mbar_wait(&bars->acc_empty[t], (fuse[t] - 1) & 1);
            fence_after_sync();
          }
          if (elect_one()) {
            const uint32_t d_addr = tmem_base + DX_ACC_COL0 + (uint32_t)t * DX_ACC_COLS;
            const uint32_t a_addr = tmem_base + (uint32_t)t * 128;
            uint64_t bd = bdesc;
            for (int ks = 0; ks < ksteps; ks += 4, bd += box_desc_step) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                if (ks + k4 < ksteps) umma_ts(d_addr, a_addr + (uint32_t)(ks + k4) * 8, bd + (uint64_t)(k4 * 2), idesc, (ks + k4) > 0 ? 1u : 0u);
              }
            }
            umma_commit(&bars->acc_full[t]);
          }
          __syncwarp();
          ++fuse[t];
        }
        if (elect_one()) {
          if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
          else umma_commit(&bars->w_empty[ws]);
        }
        __syncwarp();
        if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; bdesc = bdesc0; }
        else bdesc += slot_desc_step;
      }
      for (int t = 0; t < nt; ++t) {
        if (elect_one()) umma_commit(&bars->a_empty[t]);     // every MMA that reads this tile's dY has been issued and will complete
        __syncwarp();
        ++ause[t];
      }
    }
  } else {
    // =============================== row warps (both tiles) ===============================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t fuse[2] = {0, 0};
    int use[2] = {0, 0};
    uint32_t* myXk[2] = {sXk + ((0 * 2 + half) * NPAIR) * 128 + rl, sXk + ((1 * 2 + half) * NPAIR) * 128 + rl};
    for (int it = 0; it < p.n_iters; ++it) {
      const int nt = ntiles_of(it);
      if (nt == 0) continue;
      int64_t row[2];
      bool valid[2];
      float dxk[2][HALF];
      for (int t = 0; t < 2; ++t) {
        row[t] = tile_of(it, t) * 128 + rl;
        valid[t] = t < nt && row[t] < p.R;
      }
      // ---- stage the tiles' dY rows into TMEM (A operands); all global loads of a row half are issued before the first store
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t < nt) {
          if (use[t] > 0) {
            mbar_wait(&bars->a_empty[t], (use[t] - 1) & 1);
            fence_after_sync();
          }
          const int ncol = p.H_pad / 2;                // 32-bit columns of the A tile
          const int c_beg = half == 0 ? 0 : ((ncol / 2 + 3) & ~3);
          const int c_end = half == 0 ? ((ncol / 2 + 3) & ~3) : ncol;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(p.dyt + row[t] * p.Hs);
          uint4 buf[17];                               // (128 / 2 + 3) / 4 granules at most
          int ng = 0;
          for (int c = c_beg; c < c_end; c += 4, ++ng) {
            buf[ng] = make_uint4(0u, 0u, 0u, 0u);
            if (valid[t] && c * 2 < p.Hs) buf[ng] = *reinterpret_cast<const uint4*>(src + c);
          }
          ng = 0;
          for (int c = c_beg; c < c_end; c += 4, ++ng) {
            const uint32_t v4[4] = {buf[ng].x, buf[ng].y, buf[ng].z, buf[ng].w};
            tmem_st_x4(tmem_base + lane_addr + (uint32_t)t * 128 + c, v4);
          }
          tmem_wait_st();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->a_full[t]);
        }
      }
      // ---- this thread's X^{k-1} row halves -> private shared-memory columns; accumulators
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) dxk[t][i] = 0.f;
        if (t < nt) {
          const __nv_bfloat16* xr = p.xkt + row[t] * p.xk_pitch + half * HALF;
#pragma unroll
          for (int v8 = 0; v8 < HALF / 8; ++v8) {
            uint4 x = make_uint4(0u, 0u, 0u, 0u);
            if (valid[t] && half * HALF + v8 * 8 < p.xk_pitch) x = *reinterpret_cast<const uint4*>(xr + v8 * 8);
            myXk[t][(v8 * 4 + 0) * 128] = x.x;
            myXk[t][(v8 * 4 + 1) * 128] = x.y;
            myXk[t][(v8 * 4 + 2) * 128] = x.z;
            myXk[t][(v8 * 4 + 3) * 128] = x.w;
          }
          mbar_wait(&bars->x_full[t], use[t] & 1);
        }
      }
      // ---- fields
      for (int j = 0; j < p.m; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < nt) {
            const __nv_bfloat16* x0row = reinterpret_cast<const __nv_bfloat16*>(sX0 + (size_t)t * x0_tile) + (size_t)rl * p.mP;
            const float x0v = __bfloat162float(x0row[j]);
            mbar_wait(&bars->acc_full[t], fuse[t] & 1);
            fence_after_sync();
            const uint32_t acc = tmem_base + lane_addr + DX_ACC_COL0 + (uint32_t)t * DX_ACC_COLS + half * HALF;
            float dot = 0.f;
#pragma unroll
            for (int b0 = 0; b0 < HALF; b0 += 32) {
              constexpr int dummy = 0;
              (void)dummy;
              uint32_t v[32];
#pragma unroll
              for (int c0 = 0; c0 < 32; c0 += 8) {
                if (b0 + c0 < HALF) {
                  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                               : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                                 "=r"(v[c0 + 6]), "=r"(v[c0 + 7])
                               : "r"(acc + b0 + c0)
                               : "memory");
                }
              }
              tmem_wait_ld();
              if (b0 + 32 >= HALF) {                     // last batch: the accumulator goes back to the tensor core
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[t]);
              }
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                if (b0 + i < HALF) {
                  const uint32_t xp = myXk[t][((b0 + i) / 2) * 128];
                  const float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xp));
                  const float z0 = __uint_as_float(v[i]), z1 = __uint_as_float(v[i + 1]);
                  dxk[t][b0 + i] = fmaf(z0, x0v, dxk[t][b0 + i]);
                  dxk[t][b0 + i + 1] = fmaf(z1, x0v, dxk[t][b0 + i + 1]);
                  dot = fmaf(z0, xf.x, dot);
                  dot = fmaf(z1, xf.y, dot);
                }
              }
            }
            atomicAdd(&sDx0[((size_t)t * 128 + rl) * p.mP + j], dot);      // two addends per element (the channel halves)
            ++fuse[t];
          }
        }
      }
      // ---- tile outputs
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (valid[t]) {
          float* o = p.dxk + row[t] * p.HpQ + half * HALF;
#pragma unroll
          for (int i = 0; i < HALF; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(dxk[t][i], dxk[t][i + 1], dxk[t][i + 2], dxk[t][i + 3]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // all dX0 partials of both tiles are in shared memory
      {
        const int t = half;                              // the half-0 warps flush tile 0, the half-1 warps tile 1
        float* plane = sDx0 + ((size_t)t * 128 + rl) * p.mP;
        if (valid[t]) {
          float* gx = p.dx0 + row[t] * p.mP;
          for (int j = 0; j < p.m; ++j) gx[j] += plane[j];
        }
        for (int j = 0; j < p.m; ++j) plane[j] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // planes are zero again, x0 rows no longer needed
      __syncwarp();
      for (int t = 0; t < nt; ++t) {
        if (lane == 0) mbar_arrive(&bars->x_empty[t]);
        ++use[t];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// W fp32 [H, Hp*m] (k = i*m + j) -> W'' bf16 [m*HpQ rows (row = j*HpQ + i), HC cols (h), zero padded]
__global__ void cin_prep_wt_kernel(const float* __restrict__ W, int H, int Hp, int m, int HpQ, int HC, __nv_bfloat16* __restrict__ Wt) {
  int64_t total = (int64_t)m * HpQ * HC;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int h = (int)(e % HC);
    int64_t r = e / HC;
    int i = (int)(r % HpQ), j = (int)(r / HpQ);
    float v = 0.f;
    if (h < H && i < Hp) v = W[(int64_t)h * Hp * m + (int64_t)i * m + j];
    Wt[e] = __float2bfloat16(v);
  }
}

static int round_up_i(int a, int b) { return (a + b - 1) / b * b; }

struct CinDxGeom {
  int HpQ, H_pad, Hs, mP, HC, n_hchunks, ns;
  int n_full, tail_ks;
  size_t slot;                // bytes of one ring slot (one field, unpadded)
  size_t smem;
};

static size_t cin_dx_fixed_smem(int m, int mP, int groups) {
  return 2 * (size_t)128 * mP * 2 + (size_t)groups * 128 * (m | 1) * 4 + 8 + sizeof(CinDxBars) + 256;
}

static int cin_dx_geom(int m, int Hp, int H, int D, CinDxGeom* g) {
  if (!(D == 8 || D == 16 || D == 32 || D == 64 || D == 128) || Hp > 128 || H > 256 || m > XDFM_MAX_FIELDS) {
    xdfm_set_error("cin_bwd_dx_tc: unsupported shape (D=%d in {8..128 pow2}, Hp=%d<=128, H=%d<=256, m=%d<=64)", D, Hp, H, m);
    return XDFM_ERR_UNSUPPORTED;
  }
  g->HpQ = round_up_i(Hp, 16);
  g->H_pad = round_up_i(H, 16);
  g->Hs = round_up_i(H, 8);
  g->mP = round_up_i(m, 8);
  g->n_hchunks = (g->H_pad + 63) / 64;
  g->HC = g->n_hchunks * 64;
  g->n_full = g->H_pad / 64;
  g->tail_ks = (g->H_pad % 64) / 16;
  size_t fixed = cin_dx_fixed_smem(m, g->mP, 2);
  size_t slot = (size_t)g->HpQ * 128 * g->n_full + (size_t)g->HpQ * 32 * g->tail_ks;      // one field
  g->slot = slot;
  int ns = (int)std::min<size_t>((227 * 1024 - fixed) / slot, DX_MAX_NS * 1);
  ns = std::min(ns, DX_MAX_NS);
  if (const char* e = getenv("XDFM_DEBUG_DX_NS")) {         // profiling experiments only: shallower W'' ring
    if (atoi(e) >= 2) ns = std::min(ns, atoi(e));
  }
  if (ns < 2) {
    xdfm_set_error("cin_bwd_dx_tc: shared memory too small");
    return XDFM_ERR_UNSUPPORTED;
  }
  g->ns = ns;
  g->smem = fixed + (size_t)ns * slot;
  return XDFM_OK;
}

extern "C" int64_t xdfm_cin_bwd_dx_tc_wt_elems(int m, int Hp, int H, int D) {
  CinDxGeom g;
  if (cin_dx_geom(m, Hp, H, D, &g) != XDFM_OK) return -1;
  return (int64_t)m * g.HpQ * g.HC;
}

extern int g_cin_tc_cluster_shared;
int g_cin_dx_debug = 0;
extern "C" void xdfm_cin_dx_set_debug(int v) { g_cin_dx_debug = v; }

int g_cin_dx_pair = 0;      // measured: same time as the single-tile kernel (0.324 vs 0.326 ms, cfg2 layer 2) -> off by default
extern "C" void xdfm_cin_dx_set_pair(int v) { g_cin_dx_pair = v ? 1 : 0; }

template <int NQ>
static int launch_dx_pair(const CUtensorMap& tmW, const CinDxParams& p, size_t smem, int blocks, int cluster, cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_bwd_dx_tc_pair_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(DX_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_bwd_dx_tc_pair_kernel<NQ>, tmW, p));
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

int g_cin_dx_groups = 2;     // row warps per TMEM lane quarter in the dX kernel: 2 (default) or 4 (measured slower: 0.308 vs 0.285 ms)
extern "C" void xdfm_cin_dx_set_groups(int v) { g_cin_dx_groups = (v == 4) ? 4 : 2; }

template <int NQ, int NG>
static int launch_dx(const CUtensorMap& tmW, const CUtensorMap& tmWt, const CinDxParams& p, size_t smem, int blocks, int cluster,
                     cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_bwd_dx_tc_kernel<NQ, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3((2 + 4 * NG) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_bwd_dx_tc_kernel<NQ, NG>, tmW, tmWt, p));
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// dyt [B*D, Hs] bf16; x0t [B*D, mP] bf16; xkt rows (pitch xk_pitch) bf16; W fp32 [H, Hp*m]; wt = bf16 scratch
// [xdfm_cin_bwd_dx_tc_wt_elems]; dxk [B*D, HpQ] fp32 out (HpQ = Hp rounded up to 16); dx0 [B*D, mP] fp32 accumulated (+=).
extern "C" int xdfm_cin_bwd_dx_tc(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt,
                                  int64_t B, int m, int Hp, int H, int D, float* dxk, float* dx0, void* stream) {
  CinDxGeom g;
  int rc = cin_dx_geom(m, Hp, H, D, &g);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CHECK_ARG(((uintptr_t)dyt % 16 == 0) && ((uintptr_t)x0t % 16 == 0) && ((uintptr_t)xkt % 16 == 0) && xk_pitch % 8 == 0 &&
                     ((uintptr_t)dxk % 16 == 0),
                 "cin_bwd_dx_tc: operands must be 16-byte aligned and xk_pitch a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t total = (int64_t)m * g.HpQ * g.HC;
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64(total, 256));
    cin_prep_wt_kernel<<<blocks, 256, 0, st>>>(W, H, Hp, m, g.HpQ, g.HC, (__nv_bfloat16*)wt);
    XDFM_LAUNCH_CHECK();
  }
  int cluster = g_cin_tc_cluster_shared;
  while (cluster > 1 && ((g.HpQ / 8) % cluster) != 0) cluster >>= 1;
  CUtensorMap tmW;
  rc = xdfm_make_tmap_bf16(&tmW, wt, (uint64_t)m * g.HpQ, (uint64_t)g.HC, (uint64_t)g.HC * 2, (uint32_t)(g.HpQ / cluster), 64, 1);
  if (rc) return rc;
  // 16-wide SWIZZLE_32B boxes over the same matrix: the K tail (H_pad % 64) of every field
  CUtensorMap tmWt;
  rc = xdfm_make_tmap_bf16(&tmWt, wt, (uint64_t)m * g.HpQ, (uint64_t)g.HC, (uint64_t)g.HC * 2, (uint32_t)(g.HpQ / cluster), 16, 2);
  if (rc) return rc;
  const int64_t R = B * (int64_t)D;
  CinDxParams p;
  p.n_full = g.n_full; p.tail_ks = g.tail_ks;
  p.na_shift = g.HpQ <= 64 ? 2 : 1;
  p.dyt = (const __nv_bfloat16*)dyt; p.x0t = (const __nv_bfloat16*)x0t; p.xkt = (const __nv_bfloat16*)xkt; p.dxk = dxk; p.dx0 = dx0;
  p.R = R; p.xk_pitch = xk_pitch; p.m = m; p.mP = g.mP; p.Hp = Hp; p.HpQ = g.HpQ; p.H = H; p.H_pad = g.H_pad; p.Hs = g.Hs;
  p.n_tiles = ceil_div64(R, 128); p.n_hchunks = g.n_hchunks; p.ns = g.ns; p.debug = g_cin_dx_debug;
  int sms = xdfm_num_sms();
  int blocks = (int)std::min<int64_t>(ceil_div64(p.n_tiles, cluster) * cluster, (int64_t)(sms / cluster) * cluster);
  blocks = std::max(blocks, cluster);
  p.n_iters = (int)ceil_div64(p.n_tiles, blocks);
  if (g_cin_dx_pair && g_cin_dx_debug == 0) {
    // tile-pair kernel: needs the pair's shared-memory areas next to at least two W'' slots
    const size_t slot = (size_t)g.HpQ * 128 * g.n_hchunks;
    const size_t fixed = 2 * (size_t)128 * g.mP * 2 + (size_t)2 * 2 * (g.HpQ / 4) * 128 * 4 + 2 * (size_t)128 * g.mP * 4 +
                         sizeof(CinDxPairBars) + 256;
    int ns = (227 * 1024 > fixed) ? (int)std::min<size_t>((227 * 1024 - fixed) / slot, DX_MAX_NS) : 0;
    if (ns >= 2) {
      CinDxParams pp = p;
      pp.ns = ns;
      pp.n_iters = (int)ceil_div64(p.n_tiles, 2 * (int64_t)blocks);
      const size_t smem_pair = fixed + (size_t)ns * slot;
      switch (g.HpQ / 16) {
#define CASE_NQP(n) case n: return launch_dx_pair<n>(tmW, pp, smem_pair, blocks, cluster, st);
        CASE_NQP(1) CASE_NQP(2) CASE_NQP(3) CASE_NQP(4) CASE_NQP(5) CASE_NQP(6) CASE_NQP(7) CASE_NQP(8)
#undef CASE_NQP
      }
    }
  }
  if (g_cin_dx_groups == 4) {
    // four row warps per lane quarter: one dX0 partial plane per group in shared memory
    const size_t slot = g.slot;
    const size_t fixed4 = cin_dx_fixed_smem(m, g.mP, 4);
    int ns4 = (227 * 1024 > fixed4) ? (int)std::min<size_t>((227 * 1024 - fixed4) / slot, DX_MAX_NS) : 0;
    if (ns4 >= 2) {
      CinDxParams p4 = p;
      p4.ns = ns4;
      const size_t smem4 = fixed4 + (size_t)ns4 * slot;
      switch (g.HpQ / 16) {
#define CASE_NQ4(n) case n: return launch_dx<n, 4>(tmW, tmWt, p4, smem4, blocks, cluster, st);
        CASE_NQ4(1) CASE_NQ4(2) CASE_NQ4(3) CASE_NQ4(4) CASE_NQ4(5) CASE_NQ4(6) CASE_NQ4(7) CASE_NQ4(8)
#undef CASE_NQ4
      }
    }
  }
  switch (g.HpQ / 16) {
#define CASE_NQ(n) case n: return launch_dx<n, 2>(tmW, tmWt, p, g.smem, blocks, cluster, st);
    CASE_NQ(1) CASE_NQ(2) CASE_NQ(3) CASE_NQ(4) CASE_NQ(5) CASE_NQ(6) CASE_NQ(7) CASE_NQ(8)
#undef CASE_NQ
  }
  xdfm_set_error("cin_bwd_dx_tc: unreachable HpQ=%d", g.HpQ);
  return XDFM_ERR_UNSUPPORTED;
}
