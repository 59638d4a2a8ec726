// Shared by the dX kernels (cin_tc_bwd_dx.cu, cin_tc_bwd_dx2.cu): launch parameters and the host-side hand-over.
#pragma once
#include <stdint.h>
#include <cuda_bf16.h>

#define DX_MAX_NS 8

struct CinDxParams {
  const __nv_bfloat16* dyt;   // [R, Hs]
  const __nv_bfloat16* x0t;   // [R, mP]
  const __nv_bfloat16* xkt;   // rows with pitch xk_pitch, first Hp channels used
  float* dxk;                 // [R, HpQ] fp32 (overwritten); unused when dyp is set
  __nv_bfloat16* dyp;         // optional: dY of the layer BELOW, channels [0, HpQ) of rows with pitch dy_pitch = act'(X^{k-1}) * dXk (bf16)
  int64_t dy_pitch;
  int dy_relu;                // act' = (X^{k-1} > 0) (ReLU) or 1 (linear)
  float* dx0;                 // [2, R, mP] fp32 (overwritten): one plane per half of the X^{k-1} channels
  int64_t R, xk_pitch;
  int m, mP, Hp, HpQ, H, H_pad, Hs;
  int64_t n_tiles;
  int n_iters;
  int n_hchunks;              // 64-wide chunks of the reduction dim h per field = ceil(H_pad / 64)
  int debug;                  // diagnostic bit mask (0 in production): 1 skip contraction FMAs, 2 skip TMEM loads, 4 skip MMAs
  int ns;                     // W'' ring depth (one slot = one field GROUP on one barrier)
  // a slot holds the group's n_full 64-wide h-chunks ([fpg * HpQ rows x 128 B], SWIZZLE_128B) followed by tail_ks 16-wide chunks
  // ([fpg * HpQ rows x 32 B], SWIZZLE_32B) -- H_pad = 64 n_full + 16 tail_ks, nothing zero-padded is streamed
  int n_full, tail_ks;        // the ring keeps the full chunks of all slots first (1024-byte aligned), then the tails (256-byte aligned)
  long long* trace;           // profiling only: clock64 stamps of CTA 0's hand-offs ([tile][group][8 events]); nullptr in production
  int fpg;                    // consecutive fields contracted by one MMA group (N = fpg * HpQ <= 128)
};

bool cin_dx2_geom(int m, int HpQ, int H_pad, int mP, int* fpg_out, int* ns_out, size_t* smem_out);
int cin_dx2_launch(const void* wt, int HC, CinDxParams p, int fpg, int ns, size_t smem, int blocks, int cluster, cudaStream_t st);
