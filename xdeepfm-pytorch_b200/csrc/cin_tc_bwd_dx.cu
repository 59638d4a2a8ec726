// CIN layer backward w.r.t. the activations on the tensor cores: host side (shape checks, W'' preparation, launch hand-over).
// The kernel lives in cin_tc_bwd_dx2.cu.
//
// Replaces the dZ / dX part of torch's autograd of deepctr/layers/interaction.py:218-224 (convolution_backward input grad +
// einsum backward).  With dY = act'(y) * upstream (bf16, row layout [R, Hs], R = B*D rows r = (sample, d)):
//
//     dZ[r, (j,i)] = sum_h dY[r,h] * W[h, i*m + j]                 (implicit GEMM, never materialised)
//     dXk[r, i]    = sum_j dZ[r,(j,i)] * X0[r, j]                   (layer input gradient, fp32 row layout [R, HpQ])
//     dX0[r, j]    = sum_i dZ[r,(j,i)] * Xk[r, i]                   (two fp32 planes [2, R, mP], one per half of the channels i)
//
// W'' = weights permuted to [m*HpQ rows (j-major), H_pad cols] bf16 (K-major for this GEMM), streamed by TMA per field group.
#include "tc_common.cuh"
#include "../../include/xdfm.h"
#include "cin_tc_bwd_dx.cuh"

using namespace tc;

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
  int HpQ, H_pad, Hs, mP, HC, n_hchunks;
  int n_full, tail_ks;        // H_pad = 64 n_full + 16 tail_ks: full SWIZZLE_128B chunks and SWIZZLE_32B tail chunks of a W'' row
  int fpg, ns;                // fields per MMA group, ring depth
  size_t smem;
};

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
  if (!cin_dx2_geom(m, g->HpQ, g->H_pad, g->mP, &g->fpg, &g->ns, &g->smem)) {
    xdfm_set_error("cin_bwd_dx_tc: shape does not fit shared memory (m=%d Hp=%d H=%d)", m, Hp, H);
    return XDFM_ERR_UNSUPPORTED;
  }
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

// dyt [B*D, Hs] bf16; x0t [B*D, mP] bf16; xkt rows (pitch xk_pitch) bf16; W fp32 [H, Hp*m]; wt = bf16 scratch
// [xdfm_cin_bwd_dx_tc_wt_elems]; dxk [B*D, HpQ] fp32 out (HpQ = Hp rounded up to 16); dx0 [2, B*D, mP] fp32 out (two planes).
extern "C" int xdfm_cin_bwd_dx_tc_dy(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt,
                                     int64_t B, int m, int Hp, int H, int D, float* dxk, float* dx0, void* dy_prev, int64_t dy_pitch,
                                     int act, void* stream);

extern "C" int xdfm_cin_bwd_dx_tc(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt,
                                  int64_t B, int m, int Hp, int H, int D, float* dxk, float* dx0, void* stream) {
  return xdfm_cin_bwd_dx_tc_dy(dyt, x0t, xkt, xk_pitch, W, wt, B, m, Hp, H, D, dxk, dx0, nullptr, 0, XDFM_ACT_NONE, stream);
}

// As xdfm_cin_bwd_dx_tc; when dy_prev != NULL the kernel writes the dY rows of the layer below instead of dxk (dxk may be NULL):
// dy_prev[r, i] = act'(xkt[r, i]) * dXk[r, i] for i < HpQ, bf16, row pitch dy_pitch (>= HpQ, multiple of 8) -- valid when the layer
// below feeds ONLY this layer through those channels (split_half: channels [0, Hp) are the hidden half); act = XDFM_ACT_RELU / NONE.
extern "C" int xdfm_cin_bwd_dx_tc_dy(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt,
                                     int64_t B, int m, int Hp, int H, int D, float* dxk, float* dx0, void* dy_prev, int64_t dy_pitch,
                                     int act, void* stream) {
  CinDxGeom g;
  int rc = cin_dx_geom(m, Hp, H, D, &g);
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CHECK_ARG(((uintptr_t)dyt % 16 == 0) && ((uintptr_t)x0t % 16 == 0) && ((uintptr_t)xkt % 16 == 0) && xk_pitch % 8 == 0 &&
                     ((uintptr_t)dxk % 16 == 0) && ((uintptr_t)dx0 % 16 == 0),
                 "cin_bwd_dx_tc: operands must be 16-byte aligned and xk_pitch a multiple of 8");
  XDFM_CHECK_ARG(dy_prev != nullptr || dxk != nullptr, "cin_bwd_dx_tc: neither dxk nor dy_prev given");
  XDFM_CHECK_ARG(dy_prev == nullptr || (((uintptr_t)dy_prev % 16 == 0) && dy_pitch % 8 == 0 && dy_pitch >= g.HpQ &&
                                         (act == XDFM_ACT_RELU || act == XDFM_ACT_NONE)),
                 "cin_bwd_dx_tc: dy_prev must be 16-byte aligned with a pitch >= HpQ (multiple of 8) and a ReLU / linear activation");
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t total = (int64_t)m * g.HpQ * g.HC;
    int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64(total, 256));
    cin_prep_wt_kernel<<<blocks, 256, 0, st>>>(W, H, Hp, m, g.HpQ, g.HC, (__nv_bfloat16*)wt);
    XDFM_LAUNCH_CHECK();
  }
  int cluster = g_cin_tc_cluster_shared;
  const int64_t R = B * (int64_t)D;
  CinDxParams p = {};
  p.n_full = g.n_full; p.tail_ks = g.tail_ks;
  p.dyt = (const __nv_bfloat16*)dyt; p.x0t = (const __nv_bfloat16*)x0t; p.xkt = (const __nv_bfloat16*)xkt; p.dxk = dxk; p.dx0 = dx0;
  p.dyp = (__nv_bfloat16*)dy_prev; p.dy_pitch = dy_pitch; p.dy_relu = act == XDFM_ACT_RELU ? 1 : 0;
  p.R = R; p.xk_pitch = xk_pitch; p.m = m; p.mP = g.mP; p.Hp = Hp; p.HpQ = g.HpQ; p.H = H; p.H_pad = g.H_pad; p.Hs = g.Hs;
  p.n_tiles = ceil_div64(R, 128); p.n_hchunks = g.n_hchunks; p.debug = g_cin_dx_debug;
  const int blocks = (int)std::min<int64_t>(p.n_tiles, (int64_t)xdfm_num_sms());
  return cin_dx2_launch(wt, g.HC, p, g.fpg, g.ns, g.smem, blocks, cluster, st);
}
