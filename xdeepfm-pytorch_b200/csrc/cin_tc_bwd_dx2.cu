// CIN layer backward w.r.t. the activations on the tensor cores -- second generation of the kernel in cin_tc_bwd_dx.cu.
//
// Same mathematics (torch autograd of deepctr/layers/interaction.py:218-224, input gradients):
//
//     dZ[r, (j,i)] = sum_h dY[r,h] * W[h, i*m + j]                 (implicit GEMM, never materialised)
//     dXk[r, i]    = sum_j dZ[r,(j,i)] * X0[r, j]                   (fp32 row layout [R, HpQ])
//     dX0[r, j]    = sum_i dZ[r,(j,i)] * Xk[r, i]                   (fp32, two planes [2, R, mP]: one per channel half, summed with
//                                                                    the other layers' planes by xdfm_cin_dx0_finish)
//
// What round 1's kernel lost its time on (profiles/r01b_cin_dx_findings.md, profiles/r02_cin_dx2.md) was not the tensor core, the
// TMEM loads or the FMAs but the row warps' per-field latency chain: a dependent global load (the next tile's dY granule), the
// barrier wait, the TMEM load and the release-arrive were all serialised in front of every field's FMAs, with two warps per
// scheduler to hide them.  This version removes every long-latency operation from that loop:
//
//   * the next tile's dY reaches the row warps through TMA: a dedicated warp streams [128 rows x 64 columns] SWIZZLE_128B boxes
//     into a two-slot shared-memory ring; the row warps move one box per field from shared memory into the TMEM A buffer
//     (LDS.128 + tcgen05.st), so no global load sits in the field loop;
//   * the drain is software pipelined: the TMEM loads of batch b+1 are in flight while batch b is multiplied, and the accumulator
//     is handed back to the tensor core as soon as its last batch has been loaded;
//   * fp32 pairs are multiplied with fma.rn.f32x2, and the X^{k-1} row is kept converted to fp32 in registers (the row warps run
//     with 224 registers after setmaxnreg; the producer warpgroup gives its registers up), one third of the instructions;
//   * narrow layers (HpQ <= 64) contract FPG = 128 / HpQ consecutive fields with ONE MMA group (N = FPG * HpQ, the weight rows of
//     consecutive fields are consecutive rows of W'') and one barrier hand-off instead of one per field.
//
// Warps (384 threads): 0 = W'' stream (TMA, multicast across the cluster), 1 = MMA issuer + TMEM alloc, 2 = per-tile loads (X^0 rows,
// dY boxes), 3 = idle, 4..11 = row warps (TMEM lane quarter = warp & 3, channel half = (warp - 4) >> 2).
// TMEM columns: [0,128) / [128,256) dY of the current / next tile (A operand, TS-mode MMA), [256,384) / [384,512) two accumulators.
#include "tc_common.cuh"
#include "../../include/xdfm.h"
#include "cin_tc_bwd_dx.cuh"

using namespace tc;

#define DX2_THREADS 384
#define DX2_ACC_COL0 256
#define DX2_DY_BOX 16384      // one dY box: 128 rows x 64 bf16, SWIZZLE_128B

struct __align__(8) CinDx2Bars {
  uint64_t w_full[DX_MAX_NS], w_empty[DX_MAX_NS];
  uint64_t a_full[2], a_empty[2];      // dY tiles in TMEM
  uint64_t acc_full[2], acc_empty[2];
  uint64_t x_full[2], x_empty[2];      // X^0 rows of a tile in shared memory
  uint64_t dy_full[2], dy_empty[2];    // dY boxes in shared memory
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// two independent fp32 FMAs in one instruction (sm_100): d = a * b + c per 32-bit half
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

#define DX2_TRACE_TILES 2
#define DX2_TRACE_GROUPS 32
// stamps go to shared memory (a global store per stamp slows the stamping warp by several hundred cycles per field and the
// whole hand-off chain follows the slowest warp); CTA 0 copies them out at the end
__device__ __forceinline__ void dx2_stamp_s(uint32_t* strace, const CinDxParams& p, int it, int g, int ev) {
  if (p.trace != nullptr && it < DX2_TRACE_TILES && g < DX2_TRACE_GROUPS) {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    strace[(it * DX2_TRACE_GROUPS + g) * 16 + ev] = c;
  }
}
__device__ __forceinline__ void dx2_stamp(const CinDxParams& p, int it, int g, int ev) {
  if (p.trace != nullptr && !(p.debug & 256) && blockIdx.x == 0 && it < DX2_TRACE_TILES && g < DX2_TRACE_GROUPS)
    p.trace[((size_t)it * DX2_TRACE_GROUPS + g) * 16 + ev] = clock64();
}
// debug bit 256: the eight stamps follow ONE row warp (warp 4) through a field (NBF == 2 shapes) instead of the hand-offs
__device__ __forceinline__ void dx2_stamp_row(const CinDxParams& p, int it, int j, int ev) {
  if (p.trace != nullptr && (p.debug & 256) && blockIdx.x == 0 && threadIdx.x == 128 && it < DX2_TRACE_TILES && j < DX2_TRACE_GROUPS)
    p.trace[((size_t)it * DX2_TRACE_GROUPS + j) * 16 + ev] = clock64();
}

template <int BS>
__device__ __forceinline__ void tmem_ld_batch(uint32_t taddr, uint32_t (&v)[BS]) {
#pragma unroll
  for (int c = 0; c + 16 <= BS; c += 16) tmem_ld_x16(taddr + c, &v[c]);
  if constexpr ((BS % 16) >= 8) {
    constexpr int c0 = BS / 16 * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]), "=r"(v[c0 + 6]),
                   "=r"(v[c0 + 7])
                 : "r"(taddr + c0)
                 : "memory");
  }
  if constexpr ((BS % 8) == 4) {
    constexpr int c0 = BS - 4;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3])
                 : "r"(taddr + c0)
                 : "memory");
  }
}

// the registers of an asynchronous TMEM load may be read only after tcgen05.wait::ld: tie them to this point of the program
template <int BS>
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&v)[BS]) {
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < BS; ++i) asm volatile("" : "+r"(v[i]));
}

// mbarrier by shared-memory address, without the printf of tc::mbar_wait: the row warps' loop stays small; a protocol bug traps
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0, ok;
  do {
    // suspend-time hint: the thread sleeps in hardware until the phase completes (or 20 us pass) instead of coming back to poll
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(20000u)
                 : "memory");
    if (!ok && ++spins > (1u << 26)) {
      printf("xdfm: dX mbarrier wait timed out (block %d thread %d bar@%u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  } while (!ok);
}
// producer-side wait: the W'' / dY / X^0 producers are whole fields ahead of their consumers, so they can afford to sleep between
// polls -- a polling warp takes issue slots from the two row warps on its scheduler (clock stamps, r02q: the row warps next to the
// W'' producer handed their accumulators back ~1000 cycles per field later than those next to an idle warp)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* barp, uint32_t parity) {
  const uint32_t bar = smem_u32(barp);
  uint32_t spins = 0, ok;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(20000u)
                 : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) {
      printf("xdfm: dX producer wait timed out (block %d thread %d bar@%u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// every lane polls (measured: one polling lane + __syncwarp is slower, 0.302 vs 0.220 ms on the cfg2 wide layer)
__device__ __forceinline__ void mbar_wait_w(uint32_t bar, uint32_t parity, bool lead) {
  (void)lead;
  mbar_wait_a(bar, parity);
}
// "this buffer may be overwritten": the arriving thread only READ the buffer and those reads have completed (their values were
// used; TMEM loads: tcgen05.wait::ld + fence), so the arrive needs no release ordering.  The default (release) form also waits for
// every other memory operation the thread has in flight -- a prefetched global load or the tile's output stores, a microsecond each.
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.relaxed.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

// NQ = HpQ / 16.  Each row warp drains HALF = HpQ / 2 channels of every field in NBF batches of BS columns.
// DBG = true compiles the switch-off experiments (p.debug bits) and the clock stamps (p.trace) in; production launches DBG = false.
template <int NQ, bool DBG>
__global__ void __launch_bounds__(DX2_THREADS, 1)
cin_bwd_dx_tc2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWt,
                      const __grid_constant__ CUtensorMap tmDy, CinDxParams p) {
  constexpr int HpQ = NQ * 16;
  constexpr int HALF = HpQ / 2;
  constexpr int NBF = HALF > 32 ? 2 : 1;           // batches per field
  constexpr int BS = HALF / NBF;                   // columns per batch (multiple of 4)
  static_assert(BS * NBF == HALF && BS % 4 == 0, "dZ batch split");
  constexpr bool LOAD_ALL = HALF <= 32;            // the whole field half in one batch (measured at HALF = 56: spills, 0.198 vs 0.192 ms)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NG = p.fpg * HpQ;                                                // accumulator columns / weight rows of a field group
  const uint32_t w_box_bytes = (uint32_t)NG * 128;                           // one 64-wide h-chunk of one field group
  const uint32_t w_tail_bytes = (uint32_t)NG * 32;                           // one 16-wide h-chunk of the tail
  const uint32_t full_stride = w_box_bytes * (uint32_t)p.n_full;
  const uint32_t tail_stride = w_tail_bytes * (uint32_t)p.tail_ks;
  const uint32_t w_slot_bytes = full_stride + tail_stride;
  uint8_t* sW = smem;                                                         // ns x full chunks
  uint8_t* sWt = sW + (size_t)p.ns * full_stride;                             // ns x tail chunks
  uint8_t* sDY = smem + (((size_t)p.ns * w_slot_bytes + 1023) & ~(size_t)1023);   // 2 dY boxes
  const uint32_t x0_tile = (uint32_t)128 * p.mP * 2;
  uint8_t* sX0 = sDY + 2 * DX2_DY_BOX;                                        // 2 x [128][mP] bf16
  const int dpitch = p.m | 1;                                                 // odd pitch: a warp's 32 rows hit 32 different banks
  float* sDx0 = reinterpret_cast<float*>(sX0 + 2 * (size_t)x0_tile);          // [2 halves][128][dpitch] fp32: each thread's own dX0 row
  CinDx2Bars* bars = reinterpret_cast<CinDx2Bars*>(sDx0 + ((2 * 128 * dpitch + 1) & ~1));
  uint32_t* strace = reinterpret_cast<uint32_t*>(bars + 1);                   // DBG builds only: DX2_TRACE_TILES x GROUPS x 16 stamps
  if constexpr (DBG) {
    for (int i = threadIdx.x; i < DX2_TRACE_TILES * DX2_TRACE_GROUPS * 16; i += blockDim.x) strace[i] = 0u;
  }

  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < DX_MAX_NS; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], csize); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 8);    mbar_init(&bars->a_empty[i], 1);
      mbar_init(&bars->acc_full[i], 1);  mbar_init(&bars->acc_empty[i], 8);
      mbar_init(&bars->x_full[i], 1);    mbar_init(&bars->x_empty[i], 8);
      mbar_init(&bars->dy_full[i], 1);   mbar_init(&bars->dy_empty[i], 8);
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
  const int n_groups = (p.m + p.fpg - 1) / p.fpg;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // =============================== W'' stream ===============================
      if (lane == 0) {
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmWt);
        const int slice = NG / (int)csize;             // rows of a slot loaded (and multicast) by this CTA
        const int wr0 = (int)crank * slice;
        uint32_t ws = 0, wphase = 1;
        bool first_pass = true;
        for (int it = 0; it < p.n_iters; ++it) {
          for (int g = 0; g < n_groups; ++g) {
            if (!first_pass) mbar_wait_sleep(&bars->w_empty[ws], wphase);
            if constexpr (DBG) {
              dx2_stamp_s(strace, p, it, g, 3);
              if (p.debug & 8) {                      // experiment: no weight stream (barrier hand-offs only)
                mbar_arrive(&bars->w_full[ws]);
                if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; first_pass = false; }
                continue;
              }
            }
            mbar_arrive_expect_tx(&bars->w_full[ws], w_slot_bytes);
            for (int c = 0; c < p.n_full; ++c) {
              uint8_t* dst = sW + (size_t)ws * full_stride + (size_t)c * w_box_bytes + (size_t)wr0 * 128;
              if (csize > 1) tma_load_2d_mcast(dst, &tmW, c * 64, g * NG + wr0, &bars->w_full[ws], cmask);
              else tma_load_2d(dst, &tmW, c * 64, g * NG + wr0, &bars->w_full[ws]);
            }
            for (int t = 0; t < p.tail_ks; ++t) {
              uint8_t* dst = sWt + (size_t)ws * tail_stride + (size_t)t * w_tail_bytes + (size_t)wr0 * 32;
              if (csize > 1) tma_load_2d_mcast(dst, &tmWt, p.n_full * 64 + t * 16, g * NG + wr0, &bars->w_full[ws], cmask);
              else tma_load_2d(dst, &tmWt, p.n_full * 64 + t * 16, g * NG + wr0, &bars->w_full[ws]);
            }
            if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; first_pass = false; }
          }
        }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer (warp-uniform loop, elected lane issues) ===============================
      // (Two issuing warps taking the groups in turn measured 8 % faster on the cfg2 wide layer -- one warp's barrier waits and
      // commits overlap the other's MMAs -- but the full training flow then failed intermittently with a launch failure; one
      // issuing thread per CTA is what every reference pipeline does, so that is what ships.)
      const uint32_t idesc = make_idesc_bf16(128, NG);
      const uint64_t bdesc0 = make_desc_k_sw128(smem_u32(sW));
      const uint64_t tdesc0 = make_desc_k_sw32(smem_u32(sWt));
      const uint32_t slot_desc_step = full_stride >> 4;
      const uint32_t tslot_desc_step = tail_stride >> 4;
      const uint32_t box_desc_step = w_box_bytes >> 4;
      const uint32_t tail_desc_step = w_tail_bytes >> 4;
      uint32_t ws = 0, wphase = 0;      // ring slot / phase of the group under the cursor (every group, whoever issues it)
      uint32_t G = 0;                   // running index of the group under the cursor
      uint32_t gc = 0;                  // active groups before the cursor (accumulator = gc & 1)
      int at = 0;
      const int ksteps = p.n_full * 4;
      for (int it = 0; it < p.n_iters; ++it) {
        const bool active = tile_of(it) < p.n_tiles;
        const uint32_t abuf = (uint32_t)(at & 1);
        const uint32_t a_addr0 = tmem_base + abuf * 128;
        bool a_ready = false;
        for (int g = 0; g < n_groups; ++g, ++G) {
          {
            const uint32_t ab = gc & 1;
            if (active && !a_ready) {
              mbar_wait_w(smem_u32(&bars->a_full[abuf]), (at >> 1) & 1, lane == 0);
              fence_after_sync();
              a_ready = true;
            }
            if (active && gc >= 2) {
              mbar_wait_w(smem_u32(&bars->acc_empty[ab]), ((gc >> 1) - 1) & 1, lane == 0);
              fence_after_sync();
            }
            if constexpr (DBG) { if (lane == 0) dx2_stamp_s(strace, p, it, g, 0); }
            mbar_wait_w(smem_u32(&bars->w_full[ws]), wphase, lane == 0);
            fence_after_sync();
            if constexpr (DBG) { if (lane == 0) dx2_stamp_s(strace, p, it, g, 1); }
            if (elect_one()) {
              if (active && !(DBG && (p.debug & 4))) {
                const uint32_t d_addr = tmem_base + DX2_ACC_COL0 + ab * 128;
                uint64_t bd = bdesc0 + (uint64_t)ws * slot_desc_step;
                const uint64_t td = tdesc0 + (uint64_t)ws * tslot_desc_step;
                for (int ks = 0; ks < ksteps; ks += 4, bd += box_desc_step) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    umma_ts(d_addr, a_addr0 + (uint32_t)(ks + k4) * 8, bd + (uint64_t)(k4 * 2), idesc, (ks + k4) > 0 ? 1u : 0u);
                }
                for (int t = 0; t < p.tail_ks; ++t)
                  umma_ts(d_addr, a_addr0 + (uint32_t)(ksteps + t) * 8, td + (uint64_t)t * tail_desc_step, idesc, (ksteps + t) > 0 ? 1u : 0u);
              }
              if (csize > 1) umma_commit_mcast(&bars->w_empty[ws], cmask);
              else umma_commit(&bars->w_empty[ws]);
              if (active) umma_commit(&bars->acc_full[ab]);
            }
            __syncwarp();
            if constexpr (DBG) { if (lane == 0) dx2_stamp_s(strace, p, it, g, 2); }
          }
          if (++ws == (uint32_t)p.ns) { ws = 0; wphase ^= 1; }
          if (active) ++gc;
        }
        if (active) {
          // every MMA that reads the tile's dY has been issued and will complete
          if (elect_one()) umma_commit(&bars->a_empty[abuf]);
          __syncwarp();
          ++at;
        }
      }
    } else if (warp == 2) {
      // =============================== per-tile loads: X^0 rows and dY boxes ===============================
      if (lane == 0) {
        prefetch_tmap(&tmDy);
        uint32_t xit = 0, dyc = 0;
        for (int it = 0; it < p.n_iters; ++it) {
          const int64_t tile = tile_of(it);
          if (tile >= p.n_tiles) break;
          const uint32_t buf = xit & 1;
          if (xit >= 2) mbar_wait_sleep(&bars->x_empty[buf], ((xit >> 1) - 1) & 1);
          const int64_t r0 = tile * 128;
          const uint32_t nrows = (uint32_t)min((int64_t)128, p.R - r0);
          mbar_arrive_expect_tx(&bars->x_full[buf], nrows * (uint32_t)(p.mP * 2));
          bulk_load_1d(sX0 + (size_t)buf * x0_tile, p.x0t + r0 * p.mP, nrows * (uint32_t)(p.mP * 2), &bars->x_full[buf]);
          ++xit;
          for (int c = 0; c < p.n_hchunks; ++c, ++dyc) {
            const uint32_t slot = dyc & 1;
            if (dyc >= 2) mbar_wait_sleep(&bars->dy_empty[slot], ((dyc >> 1) - 1) & 1);
            mbar_arrive_expect_tx(&bars->dy_full[slot], DX2_DY_BOX);
            tma_load_2d(sDY + (size_t)slot * DX2_DY_BOX, &tmDy, c * 64, (int)r0, &bars->dy_full[slot]);   // rows / columns past the end: zeros
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =============================== row warps ===============================
    // The field loop below is a chain of short dependent operations with two warps per scheduler to hide them: every instruction
    // that is not a TMEM load or an FMA costs as much as one that is (ncu, r02o: 330 instructions per field and warp, 56 of them
    // FMAs).  So: addresses and parities live in registers (opaque to the compiler, which otherwise re-derives them from
    // threadIdx inside the loop), one barrier wait and one arrive per field GROUP, no per-field bookkeeping beyond two adds.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;              // channel half: channels [half * HALF, (half + 1) * HALF) of every field
    const int rl = q * 32 + lane;
    const bool lead = lane == 0;
    const int n_gran = p.H_pad / 8;                // 16-byte granules of a dY row that the MMAs read
    const int n_boxes = p.n_hchunks;
    const int m = p.m, fpg = p.fpg;
    auto keep = [](uint32_t v) { asm volatile("" : "+r"(v)); return v; };   // value the compiler must hold in a register
    const uint32_t bar_acc_full = keep(smem_u32(&bars->acc_full[0])), bar_acc_empty = keep(smem_u32(&bars->acc_empty[0]));
    const uint32_t bar_dy_full = keep(smem_u32(&bars->dy_full[0])), bar_dy_empty = keep(smem_u32(&bars->dy_empty[0]));
    const uint32_t a_taddr = keep(tmem_base + ((uint32_t)(q * 32) << 16));
    const uint32_t acc_taddr = keep(a_taddr + DX2_ACC_COL0 + (uint32_t)(half * HALF));
    const uint32_t dy_row = keep(smem_u32(sDY) + (uint32_t)rl * 128);
    const uint32_t dy_swz = keep((uint32_t)(rl & 7) << 4);
    const uint32_t plane0 = keep(smem_u32(sDx0) + (uint32_t)((half * 128 + rl) * dpitch) * 4);
    uint32_t gc = 0;                               // field groups drained so far
    uint32_t dyc = 0;                              // dY boxes consumed so far
    int at = 0;
    // one dY box of tile buffer `tb`: shared memory -> TMEM; the two channel halves split the box's eight granules
    auto stage_box = [&](uint32_t tb, int c) {
      const uint32_t slot = dyc & 1;
      mbar_wait_a(bar_dy_full + slot * 8, (dyc >> 1) & 1);
      uint32_t gv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t g16 = (uint32_t)(half * 4 + u) << 4;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(gv[u][0]), "=r"(gv[u][1]), "=r"(gv[u][2]), "=r"(gv[u][3])
                     : "r"(dy_row + slot * DX2_DY_BOX + (g16 ^ dy_swz)));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int g = half * 4 + u;
        if (c * 8 + g < n_gran && !(DBG && (p.debug & 32))) tmem_st_x4(a_taddr + tb * 128 + (uint32_t)(c * 32 + g * 4), gv[u]);
      }
      __syncwarp();
      if (lead) mbar_arrive_a(bar_dy_empty + slot * 8);
      ++dyc;
    };
    auto publish_a = [&](uint32_t tb) {
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lead) mbar_arrive(&bars->a_full[tb]);
    };
    // this thread's X^{k-1} row half, raw bf16 (4 channels per uint2): loaded for the first tile here, for every later tile while the
    // previous tile's last fields are drained
    uint2 xraw[HALF / 4];
    auto load_xk = [&](int64_t row, bool ok) {
      const __nv_bfloat16* xr = p.xkt + row * p.xk_pitch + half * HALF;
#pragma unroll
      for (int v4 = 0; v4 < HALF / 4; ++v4) {
        xraw[v4] = make_uint2(0u, 0u);
        if (ok && half * HALF + v4 * 4 < p.xk_pitch && !(DBG && (p.debug & 64))) xraw[v4] = *reinterpret_cast<const uint2*>(xr + v4 * 4);
      }
    };
    for (int it = 0; it < p.n_iters; ++it) {
      const int64_t tile = tile_of(it);
      if (tile >= p.n_tiles) break;
      const int64_t row = tile * 128 + rl;
      const bool valid = row < p.R;
      const uint32_t abuf = (uint32_t)(at & 1), nbuf = abuf ^ 1u;
      if (at == 0) {                                 // the CTA's first tile: nothing to hide the staging behind
        for (int c = 0; c < n_boxes; ++c) stage_box(abuf, c);
        publish_a(abuf);
        load_xk(row, valid);
      }
      const bool has_next = (it + 1 < p.n_iters) && tile_of(it + 1) < p.n_tiles;
      const int64_t nrow = tile_of(it + 1) * 128 + rl;
      int sc = has_next ? 0 : n_boxes;               // boxes of the next tile staged so far
      if (has_next && at >= 1) {                     // the other A buffer was read by the previous tile's MMAs: long complete
        mbar_wait_a(smem_u32(&bars->a_empty[nbuf]), (((at + 1) >> 1) - 1) & 1);
        fence_after_sync();
      }
      const int stage_from = max(1, m / 2 - 2);      // the next tile's boxes move during the middle fields of this tile
      int xk_at = has_next ? max(0, m - 3) : m;      // ... and its X^{k-1} row is requested three fields before the end
      // ---- this row's operands
      const int buf = at & 1;
      uint64_t xkf[HALF / 2];                        // X^{k-1} row half as fp32 pairs
#pragma unroll
      for (int v4 = 0; v4 < HALF / 4; ++v4) {
        xkf[v4 * 2 + 0] = pack2(xraw[v4].x << 16, xraw[v4].x & 0xffff0000u);
        xkf[v4 * 2 + 1] = pack2(xraw[v4].y << 16, xraw[v4].y & 0xffff0000u);
      }
      mbar_wait_a(smem_u32(&bars->x_full[buf]), (at >> 1) & 1);
      uint32_t x0a = smem_u32(sX0) + (uint32_t)buf * x0_tile + (uint32_t)(rl * p.mP * 2);   // &x0[r, j] (bf16), advanced per field
      uint32_t pla = plane0;                                                                 // &dX0 partial [r, j], advanced per field
      uint64_t dxk[HALF / 2];
#pragma unroll
      for (int i = 0; i < HALF / 2; ++i) dxk[i] = 0ull;
      uint32_t xb;                                   // x0[r, j] of the field being drained, fetched one field ahead
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(xb) : "r"(x0a));

      // ---- the drain: per field group one wait, per field NBF x (TMEM load, wait, FMAs), the accumulator goes back to the tensor
      // core as soon as its last columns are in registers
      for (int j0 = 0; j0 < m; j0 += fpg) {
        const uint32_t ab = gc & 1;
        mbar_wait_a(bar_acc_full + ab * 8, (gc >> 1) & 1);
        fence_after_sync();
        const int nf = min(fpg, m - j0);
        uint32_t taddr = acc_taddr + ab * 128;
        for (int f = 0; f < nf; ++f, taddr += HpQ) {
          const uint32_t xcur = xb << 16;
          x0a += 2;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(xb) : "r"(x0a));        // next field's x0 (one past the row's end at the last field: in bounds)
          const uint64_t x0p = pack2(xcur, xcur);
          uint64_t dacc[2] = {0ull, 0ull};
          if constexpr (LOAD_ALL) {
            // the whole field half in registers first: the accumulator goes back to the tensor core before the first FMA (the
            // hand-off chain MMA -> rows -> MMA sets the pace; FMAs inside it cost a field time each)
            uint32_t v[HALF];
            if (DBG && (p.debug & 2)) {
#pragma unroll
              for (int i = 0; i < HALF; ++i) v[i] = 0u;
            } else {
              tmem_ld_batch<HALF>(taddr, v);
            }
            tmem_ld_fence<HALF>(v);
            if (f == nf - 1) {
              fence_before_sync();
              __syncwarp();
              if (lead) mbar_arrive_a(bar_acc_empty + ab * 8);
              if constexpr (DBG) { if (lead) dx2_stamp_s(strace, p, it, j0 / fpg, 4 + warp); }
            }
            if (!(DBG && (p.debug & 1))) {
#pragma unroll
              for (int i = 0; i < HALF; i += 2) {
                const uint64_t z = pack2(v[i], v[i + 1]);
                dxk[i / 2] = ffma2(z, x0p, dxk[i / 2]);
                dacc[(i / 2) & 1] = ffma2(z, xkf[i / 2], dacc[(i / 2) & 1]);
              }
            }
          } else {
#pragma unroll
          for (int nb = 0; nb < NBF; ++nb) {
            uint32_t v[BS];
            if (DBG && (p.debug & 2)) {
#pragma unroll
              for (int i = 0; i < BS; ++i) v[i] = 0u;
            } else {
              tmem_ld_batch<BS>(taddr + (uint32_t)(nb * BS), v);
            }
            tmem_ld_fence<BS>(v);
            if (nb == NBF - 1 && f == nf - 1) {      // the group's last columns are in registers
              fence_before_sync();
              __syncwarp();
              if (lead) mbar_arrive_a(bar_acc_empty + ab * 8);
              if constexpr (DBG) { if (lead) dx2_stamp_s(strace, p, it, j0 / fpg, 4 + warp); }
            }
            if (!(DBG && (p.debug & 1))) {
#pragma unroll
              for (int i = 0; i < BS; i += 2) {
                const int ci = (nb * BS + i) / 2;
                const uint64_t z = pack2(v[i], v[i + 1]);
                dxk[ci] = ffma2(z, x0p, dxk[ci]);
                dacc[(i / 2) & 1] = ffma2(z, xkf[ci], dacc[(i / 2) & 1]);
              }
            }
          }
          }
          float a0, a1, b0, b1;
          unpack2(dacc[0], a0, a1);
          unpack2(dacc[1], b0, b1);
          // dX0[r, j] partial of this thread's channel half -> its own shared-memory row (written out at tile end)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(pla), "f"((a0 + a1) + (b0 + b1)) : "memory");
          pla += 4;
        }
        ++gc;
        const int jn = j0 + nf;                      // fields drained so far
        if (sc < n_boxes && jn > stage_from) {       // one dY box of the next tile per group
          stage_box(nbuf, sc);
          if (++sc == n_boxes) publish_a(nbuf);
        }
        if (jn > xk_at) {
          load_xk(nrow, nrow < p.R);
          xk_at = m;
        }
      }
      while (sc < n_boxes) {                         // shapes with few groups: the rest of the next tile's dY
        stage_box(nbuf, sc);
        if (++sc == n_boxes) publish_a(nbuf);
      }
      // ---- tile outputs: plain stores, nothing the next tile has to wait for
      if (valid && !(DBG && (p.debug & 16))) {
        if (p.dyp != nullptr) {
          // the layer below needs dY = act'(y) * dXk and y is this layer's X^{k-1}, which sits in registers: write its dY rows (bf16,
          // hidden-half channels) instead of fp32 dXk -- half the bytes, and the fp32 round trip through the dY kernel is gone
          uint4* o = reinterpret_cast<uint4*>(p.dyp + row * p.dy_pitch + half * HALF);
#pragma unroll
          for (int i8 = 0; i8 < HALF / 8; ++i8) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float g0, g1, y0, y1;
              unpack2(dxk[i8 * 4 + u], g0, g1);
              unpack2(xkf[i8 * 4 + u], y0, y1);
              if (p.dy_relu) {
                g0 = y0 > 0.f ? g0 : 0.f;
                g1 = y1 > 0.f ? g1 : 0.f;
              }
              const __nv_bfloat162 t = __floats2bfloat162_rn(g0, g1);
              w[u] = *reinterpret_cast<const uint32_t*>(&t);
            }
            o[i8] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
        float* o = p.dxk + row * p.HpQ + half * HALF;
#pragma unroll
        for (int i = 0; i < HALF / 2; i += 2) {
          float4 f;
          unpack2(dxk[i], f.x, f.y);
          unpack2(dxk[i + 1], f.z, f.w);
          *reinterpret_cast<float4*>(o + 2 * i) = f;
        }
        }
        // this thread's dX0 partials (its own shared-memory row) -> plane `half` of dx0
        const float* myplane = sDx0 + (size_t)(half * 128 + rl) * dpitch;
        float4* g4 = reinterpret_cast<float4*>(p.dx0 + ((int64_t)half * p.R + row) * p.mP);
        for (int i = 0; i < p.mP / 4; ++i) {
          float a[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) a[t] = (i * 4 + t < m) ? myplane[i * 4 + t] : 0.f;
          g4[i] = make_float4(a[0], a[1], a[2], a[3]);
        }
      }
      __syncwarp();
      if (lead) mbar_arrive_a(smem_u32(&bars->x_empty[buf]));
      ++at;
    }
  }
  fence_before_sync();
  __syncthreads();
  if constexpr (DBG) {
    if (p.trace != nullptr && blockIdx.x == 0)
      for (int i = threadIdx.x; i < DX2_TRACE_TILES * DX2_TRACE_GROUPS * 16; i += blockDim.x) p.trace[i] = (long long)strace[i];
  }
  if (csize > 1) cluster_sync_all();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t cin_dx2_fixed_smem(int m, int mP) {
  return 1024 + 2 * (size_t)DX2_DY_BOX + 2 * (size_t)128 * mP * 2 + (size_t)2 * 128 * (m | 1) * 4 + 8 + sizeof(CinDx2Bars) + 256;
}

// geometry of the second-generation kernel: fields per group, ring depth, shared memory; false when the shape does not fit
bool cin_dx2_geom(int m, int HpQ, int H_pad, int mP, int* fpg_out, int* ns_out, size_t* smem_out) {
  const int n_full = H_pad / 64, tail_ks = (H_pad % 64) / 16;
  const size_t fixed = cin_dx2_fixed_smem(m, mP);
  const size_t budget = 227 * 1024;
  if (fixed >= budget) return false;
  int fpg = std::max(1, std::min(128 / HpQ, m));
  int ns_cap = DX_MAX_NS;
  if (const char* e = getenv("XDFM_DEBUG_DX_NS")) {          // profiling experiments only: shallower W'' ring, fewer fields per group
    if (atoi(e) >= 2) ns_cap = std::min(ns_cap, atoi(e));
  }
  if (const char* e = getenv("XDFM_DEBUG_DX_FPG")) {
    if (atoi(e) >= 1) fpg = std::min(fpg, atoi(e));
  }
  for (;; --fpg) {
    const size_t slot = (size_t)fpg * HpQ * (128 * n_full + 32 * tail_ks);
    const int ns = (int)std::min<size_t>((budget - fixed) / slot, ns_cap);
    if (ns >= 3 || (fpg == 1 && ns >= 2)) {
      *fpg_out = fpg;
      *ns_out = ns;
      *smem_out = fixed + (size_t)ns * slot;
      return true;
    }
    if (fpg == 1) return false;
  }
}

template <int NQ, bool DBG>
static int launch_dx2(const CUtensorMap& tmW, const CUtensorMap& tmWt, const CUtensorMap& tmDy, const CinDxParams& p, size_t smem,
                      int blocks, int cluster, cudaStream_t st) {
  XDFM_CUDA(cudaFuncSetAttribute(cin_bwd_dx_tc2_kernel<NQ, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(DX2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XDFM_CUDA(cudaLaunchKernelEx(&cfg, cin_bwd_dx_tc2_kernel<NQ, DBG>, tmW, tmWt, tmDy, p));
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

long long* g_cin_dx_trace = nullptr;
// profiling only: device buffer of DX2_TRACE_TILES * DX2_TRACE_GROUPS * 16 int64 that CTA 0 of the next dX launches stamps (nullptr = off)
extern "C" void xdfm_cin_dx_set_trace(void* buf) { g_cin_dx_trace = (long long*)buf; }

// called by xdfm_cin_bwd_dx_tc (cin_tc_bwd_dx.cu) after W'' has been written to `wt`; p carries the shape, the operands and the tile schedule
int cin_dx2_launch(const void* wt, int HC, CinDxParams p, int fpg, int ns, size_t smem, int blocks, int cluster, cudaStream_t st) {
  const int NG = fpg * p.HpQ;
  while (cluster > 1 && ((NG / 8) % cluster) != 0) cluster >>= 1;
  CUtensorMap tmW, tmWt, tmDy;
  int rc = xdfm_make_tmap_bf16(&tmW, wt, (uint64_t)p.m * p.HpQ, (uint64_t)HC, (uint64_t)HC * 2, (uint32_t)(NG / cluster), 64, 1);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmWt, wt, (uint64_t)p.m * p.HpQ, (uint64_t)HC, (uint64_t)HC * 2, (uint32_t)(NG / cluster), 16, 2);
  if (rc) return rc;
  rc = xdfm_make_tmap_bf16(&tmDy, p.dyt, (uint64_t)p.R, (uint64_t)p.Hs, (uint64_t)p.Hs * 2, 128, 64, 1);
  if (rc) return rc;
  p.fpg = fpg;
  p.ns = ns;
  p.trace = g_cin_dx_trace;
  blocks = std::max(blocks / cluster * cluster, cluster);
  p.n_iters = (int)ceil_div64(p.n_tiles, blocks);
  if (p.debug != 0 || p.trace != nullptr) {       // profiling builds of the two cfg2 shapes only (+ 4 KB of stamps in shared memory)
    const size_t smem_dbg = smem + DX2_TRACE_TILES * DX2_TRACE_GROUPS * 16 * 4;
    if (smem_dbg <= 227 * 1024) {
      if (p.HpQ == 112) return launch_dx2<7, true>(tmW, tmWt, tmDy, p, smem_dbg, blocks, cluster, st);
      if (p.HpQ == 32) return launch_dx2<2, true>(tmW, tmWt, tmDy, p, smem_dbg, blocks, cluster, st);
    }
    p.debug = 0;
  }
  switch (p.HpQ / 16) {
#define CASE_NQ(n) case n: return launch_dx2<n, false>(tmW, tmWt, tmDy, p, smem, blocks, cluster, st);
    CASE_NQ(1) CASE_NQ(2) CASE_NQ(3) CASE_NQ(4) CASE_NQ(5) CASE_NQ(6) CASE_NQ(7) CASE_NQ(8)
#undef CASE_NQ
  }
  xdfm_set_error("cin_bwd_dx_tc: unreachable HpQ=%d", p.HpQ);
  return XDFM_ERR_UNSUPPORTED;
}
