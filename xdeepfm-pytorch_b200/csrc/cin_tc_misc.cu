// Elementwise / layout helpers of the tensor-core CIN path (row layout: one row per (sample, d), channels contiguous).
#include "common.cuh"
#include "../../include/xdfm.h"

// dY[r, h] = act'(y[r,h]) * ( direct-path grad (channels [hdb, H)) + next-layer grad (channels [0, n_next)) ), bf16 row layout.
// One thread = 8 consecutive channels of one row.
__global__ void __launch_bounds__(256) cin_dy_rows_kernel(const __nv_bfloat16* __restrict__ yt, int64_t R, int D, int H, int Hs, int hdb,
                                                          const float* __restrict__ dpooled, const float* __restrict__ dmaps, int fm_total,
                                                          int col_off, const float* __restrict__ dnext, int64_t dnext_pitch, int n_next,
                                                          int act, __nv_bfloat16* __restrict__ dyt) {
  const int g8 = Hs / 8;
  const int64_t total = R * g8;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / g8;
    const int h0 = (int)(e - r * g8) * 8;
    const int64_t b = r / D;
    const int d = (int)(r - b * D);
    const uint4 yv = *reinterpret_cast<const uint4*>(yt + r * Hs + h0);
    const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(&yv);
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int h = h0 + i;
      float v = 0.f;
      if (h < H) {
        if (h >= hdb) {
          if (dpooled) v += dpooled[b * fm_total + col_off + (h - hdb)];
          if (dmaps) v += dmaps[(b * fm_total + col_off + (h - hdb)) * (int64_t)D + d];
        }
        if (h < n_next && dnext) v += dnext[r * dnext_pitch + h];
        if (act == XDFM_ACT_RELU && !(__bfloat162float(yb[i]) > 0.f)) v = 0.f;
      }
      g[i] = v;
    }
    uint4 o;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(g[0], g[1]), t1 = __floats2bfloat162_rn(g[2], g[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(g[4], g[5]), t3 = __floats2bfloat162_rn(g[6], g[7]);
    o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
    o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
    *reinterpret_cast<uint4*>(dyt + r * Hs + h0) = o;
  }
}

extern "C" int xdfm_cin_dy_rows(const void* yt, int64_t B, int D, int H, int Hs, int direct_begin, const float* dpooled, const float* dmaps,
                                int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next, int act, void* dyt,
                                void* stream) {
  XDFM_CHECK_ARG(Hs % 8 == 0 && Hs >= H, "cin_dy_rows: Hs=%d must be a multiple of 8 and >= H=%d", Hs, H);
  XDFM_CHECK_ARG(act == XDFM_ACT_RELU || act == XDFM_ACT_NONE, "cin_dy_rows: activation %d not supported on the bf16 path", act);
  const int64_t R = B * (int64_t)D;
  const int64_t total = R * (Hs / 8);
  if (total == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  cin_dy_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)yt, R, D, H, Hs, direct_begin, dpooled, dmaps, fm_total,
                                                               col_off, dnext, dnext_pitch, n_next, act, (__nv_bfloat16*)dyt);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// rows [R, CP] fp32 (row layout) -> [B, C, D] fp32 (reference layout), optionally accumulating: out (+)= in
__global__ void from_rows_f32_kernel(const float* __restrict__ xt, int64_t B, int C, int D, int CP, float* __restrict__ x, int accumulate) {
  int64_t total = B * (int64_t)C * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(e % D);
    int64_t bc = e / D;
    int c = (int)(bc % C);
    int64_t b = bc / C;
    float v = xt[(b * D + d) * (int64_t)CP + c];
    x[e] = accumulate ? x[e] + v : v;
  }
}

extern "C" int xdfm_from_rows_f32(const float* xt, int64_t B, int C, int D, int CP, float* x, int accumulate, void* stream) {
  int64_t total = B * (int64_t)C * D;
  if (total == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  from_rows_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(xt, B, C, D, CP, x, accumulate);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// a[r, c] += b[r, c] for c < C on two row-layout fp32 matrices with different pitches (layer 0: dX0 += dXk)
__global__ void add_rows_f32_kernel(float* __restrict__ a, int64_t pa, const float* __restrict__ b, int64_t pb, int64_t R, int C) {
  int64_t total = R * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / C;
    int c = (int)(e - r * C);
    a[r * pa + c] += b[r * pb + c];
  }
}

extern "C" int xdfm_add_rows_f32(float* a, int64_t pitch_a, const float* b, int64_t pitch_b, int64_t R, int C, void* stream) {
  int64_t total = R * C;
  if (total == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  add_rows_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, pitch_a, b, pitch_b, R, C);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// dX0 of the whole CIN in the reference layout: x[b, c, d] = extra[r, c] + sum over the n_planes planes of parts[plane][r, c],
// r = b * D + d.  parts = the dX kernels' per-layer, per-channel-half dX0 planes ([n_planes, R, CP] fp32, summed in plane order:
// deterministic); extra = layer 0's dXk rows (X^{k-1} is X^0 itself there), pitch extra_pitch.  One block per sample-group keeps
// both the row-layout reads and the [B, C, D] writes coalesced through a shared-memory transpose.
__global__ void __launch_bounds__(256) cin_dx0_finish_kernel(const float* __restrict__ parts, int n_planes, const float* __restrict__ extra,
                                                             int64_t extra_pitch, int64_t B, int C, int D, int CP, float* __restrict__ x) {
  extern __shared__ float sh[];                        // [rows of a pass][CP + 1]
  const int64_t R = B * (int64_t)D;
  const int spb = max(1, 256 / D);                     // samples per block pass
  const int rows = spb * D;
  const int pitch = CP + 1;
  const int v4_per_row = CP / 4;
  for (int64_t b0 = (int64_t)blockIdx.x * spb; b0 < B; b0 += (int64_t)gridDim.x * spb) {
    const int64_t r0 = b0 * D;
    const int nrows = (int)min((int64_t)rows, R - r0);
    // 128-bit reads of every plane (rows of a pass are contiguous in each plane); all loads of an element group before the adds
    for (int e = threadIdx.x; e < nrows * v4_per_row; e += blockDim.x) {
      const int rr = e / v4_per_row, c4 = e - rr * v4_per_row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (extra != nullptr) v = *reinterpret_cast<const float4*>(extra + (r0 + rr) * extra_pitch + c4 * 4);
      const float4* src = reinterpret_cast<const float4*>(parts + (r0 + rr) * CP) + c4;
      const int64_t plane_v4 = R * (int64_t)CP / 4;
      int pl = 0;
      for (; pl + 2 <= n_planes; pl += 2) {
        const float4 a = src[(int64_t)pl * plane_v4], bq = src[(int64_t)(pl + 1) * plane_v4];
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        v.x += bq.x; v.y += bq.y; v.z += bq.z; v.w += bq.w;
      }
      if (pl < n_planes) {
        const float4 a = src[(int64_t)pl * plane_v4];
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      float* d = sh + rr * pitch + c4 * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    const int nsamp = nrows / D;
    for (int e = threadIdx.x; e < nsamp * C * D; e += blockDim.x) {
      const int d = e % D;
      const int sc = e / D;
      const int c = sc % C, sidx = sc / C;
      x[((b0 + sidx) * C + c) * (int64_t)D + d] = sh[(sidx * D + d) * pitch + c];
    }
    __syncthreads();
  }
}

extern "C" int xdfm_cin_dx0_finish(const float* parts, int n_planes, const float* extra, int64_t extra_pitch, int64_t B, int C, int D, int CP,
                                   float* x, void* stream) {
  XDFM_CHECK_ARG(D >= 1 && D <= 256 && CP >= C && CP % 4 == 0 && n_planes >= 0, "cin_dx0_finish: bad shape D=%d C=%d CP=%d planes=%d", D, C, CP,
                 n_planes);
  XDFM_CHECK_ARG(extra == nullptr || (extra_pitch >= CP && extra_pitch % 4 == 0 && (uintptr_t)extra % 16 == 0),
                 "cin_dx0_finish: extra rows need a pitch >= CP, a multiple of 4, 16-byte aligned");
  XDFM_CHECK_ARG((uintptr_t)parts % 16 == 0, "cin_dx0_finish: planes must be 16-byte aligned");
  if (B == 0) return XDFM_OK;
  const int spb = std::max(1, 256 / D);
  const size_t sm = (size_t)spb * D * (CP + 1) * sizeof(float);
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 6, ceil_div64(B, spb));
  if (sm > 48 * 1024) XDFM_CUDA(cudaFuncSetAttribute(cin_dx0_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  cin_dx0_finish_kernel<<<blocks, 256, sm, (cudaStream_t)stream>>>(parts, n_planes, extra, extra_pitch, B, C, D, CP, x);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// Fused form of cin_dy_rows + rows_to_cols: one pass over y / upstream gradients produces dY in BOTH layouts the backward needs --
// row layout dyt [R, Hs] (A operand of the dX kernel) and channel-major dyT [H_pad, R] (B operand of the dW kernel) -- through a
// 64 x 64 shared-memory tile.  Channels >= H are written as zeros in both.
__global__ void __launch_bounds__(256) cin_dy_rows_cols_kernel(const __nv_bfloat16* __restrict__ yt, int64_t R, int D, int H, int Hs, int H_pad,
                                                               int hdb, const float* __restrict__ dpooled, const float* __restrict__ dmaps,
                                                               int fm_total, int col_off, const float* __restrict__ dnext,
                                                               int64_t dnext_pitch, int n_next, int act, __nv_bfloat16* __restrict__ dyt,
                                                               __nv_bfloat16* __restrict__ dyT, float* __restrict__ part) {
  __shared__ __align__(16) __nv_bfloat16 tile[64][72];       // 144-byte rows: 16-byte aligned granules
  __shared__ float psum[4][64];                              // bias gradient: column sums of this tile's four 16-row groups
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const bool hidden_ready = dnext == nullptr && dnext_pitch < 0;
  {
    // One thread = one row x two granules of 8 channels.  A granule is (A) entirely inside the hidden half the dX kernel of the layer
    // above already wrote: copied as it is; (B) entirely direct-connect with a pooled upstream gradient: 2 x 128-bit loads under
    // the activation mask; (C) anything else (the granule that straddles the split, un-pooled maps, fp32 dnext, ragged tails): per
    // element.  (The per-element form for every granule cost 225 instructions per granule and made the kernel issue-bound.)
    const int rr = threadIdx.x >> 2, cg = (threadIdx.x & 3) * 16;
    const int64_t r = r0 + rr;
    const bool row_ok = r < R;
    const uint32_t b32 = row_ok ? (uint32_t)r / (uint32_t)D : 0u;          // R < 2^32: checked by the host
    const int64_t b = (int64_t)b32;
    const int d = (int)((uint32_t)r - b32 * (uint32_t)D);
    const __nv_bfloat16* yrow = yt + r * Hs;
    __nv_bfloat16* drow = dyt + r * Hs;
    const int hid_end = hidden_ready ? n_next : 0;
    const bool pooled_fast = dpooled != nullptr && dmaps == nullptr && (reinterpret_cast<uintptr_t>(dpooled) & 15) == 0;
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      const int h0 = c0 + cg + g8 * 8;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      bool store = false;
      if (row_ok && h0 < Hs) {
        const int64_t pidx = b * fm_total + col_off + (h0 - hdb);
        if (h0 + 8 <= hid_end) {
          o = *reinterpret_cast<const uint4*>(drow + h0);                                  // (A)
        } else if (pooled_fast && h0 >= hdb && h0 >= n_next && h0 + 8 <= H && (pidx & 3) == 0) {
          const uint4 yv = *reinterpret_cast<const uint4*>(yrow + h0);                     // (B)
          const float4 p0 = *reinterpret_cast<const float4*>(dpooled + pidx), p1 = *reinterpret_cast<const float4*>(dpooled + pidx + 4);
          float g[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          if (act == XDFM_ACT_RELU) {
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!(__uint_as_float(yw[i] << 16) > 0.f)) g[2 * i] = 0.f;                   // bf16 -> fp32 is a 16-bit shift
              if (!(__uint_as_float(yw[i] & 0xffff0000u) > 0.f)) g[2 * i + 1] = 0.f;
            }
          }
          __nv_bfloat162 t0 = __floats2bfloat162_rn(g[0], g[1]), t1 = __floats2bfloat162_rn(g[2], g[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(g[4], g[5]), t3 = __floats2bfloat162_rn(g[6], g[7]);
          o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
          o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
          store = true;
        } else {
          float g[8];                                                                      // (C)
          const uint4 yv = *reinterpret_cast<const uint4*>(yrow + h0);
          const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(&yv);
          // next layer's input gradient for these 8 channels: two 128-bit loads when the whole granule lies inside it
          float dn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) dn[i] = 0.f;
          // dnext_pitch < 0: the hidden-half channels [0, n_next) of dyt were already written by the dX kernel of the layer above
          // (xdfm_cin_bwd_dx_tc_dy); they are taken over as they are, the direct-connect channels are computed as usual
          uint4 old8 = make_uint4(0u, 0u, 0u, 0u);
          if (hidden_ready && h0 < n_next) old8 = *reinterpret_cast<const uint4*>(drow + h0);
          const __nv_bfloat16* oldb = reinterpret_cast<const __nv_bfloat16*>(&old8);
          if (dnext != nullptr && h0 < n_next) {
            const float* dsrc = dnext + r * dnext_pitch + h0;
            if (h0 + 8 <= n_next && (dnext_pitch & 3) == 0) {
              const float4 a = *reinterpret_cast<const float4*>(dsrc), c = *reinterpret_cast<const float4*>(dsrc + 4);
              dn[0] = a.x; dn[1] = a.y; dn[2] = a.z; dn[3] = a.w; dn[4] = c.x; dn[5] = c.y; dn[6] = c.z; dn[7] = c.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (h0 + i < n_next) dn[i] = dsrc[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int h = h0 + i;
            float v = 0.f;
            if (h < H) {
              if (h >= hdb) {
                if (dpooled) v += dpooled[pidx + i];
                if (dmaps) v += dmaps[(pidx + i) * (int64_t)D + d];
              }
              v += dn[i];
              if (act == XDFM_ACT_RELU && !(__bfloat162float(yb[i]) > 0.f)) v = 0.f;
              if (hidden_ready && h < n_next) v = __bfloat162float(oldb[i]);       // exact: re-rounded to the same bf16 below
            }
            g[i] = v;
          }
          __nv_bfloat162 t0 = __floats2bfloat162_rn(g[0], g[1]), t1 = __floats2bfloat162_rn(g[2], g[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(g[4], g[5]), t3 = __floats2bfloat162_rn(g[6], g[7]);
          o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
          o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
          store = true;
        }
      }
      *reinterpret_cast<uint4*>(&tile[rr][cg + g8 * 8]) = o;
      if (store) *reinterpret_cast<uint4*>(drow + h0) = o;
    }
  }
  __syncthreads();
  {
    // lanes walk the channel axis (consecutive 2-byte columns of a tile row: conflict-free), warps pairs the four 16-row groups
    const int cc = threadIdx.x & 63, rg = (threadIdx.x >> 6) * 16;
    const int c = c0 + cc;
    if (c < H_pad) {
      __align__(16) __nv_bfloat16 col[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) col[i] = tile[rg + i][cc];
      if (part != nullptr) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc += __bfloat162float(col[i]);       // rows past R hold zeros
        psum[threadIdx.x >> 6][cc] = sacc;
      }
      const int64_t r = r0 + rg;
      __nv_bfloat16* dst = dyT + (int64_t)c * R + r;
      if (r + 16 <= R && ((R & 7) == 0)) {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&col[0]);
        *reinterpret_cast<uint4*>(dst + 8) = *reinterpret_cast<const uint4*>(&col[8]);
      } else {
        for (int i = 0; i < 16; ++i)
          if (r + i < R) dst[i] = col[i];
      }
    }
  }
  if (part != nullptr) {
    __syncthreads();
    const int c = c0 + (int)threadIdx.x;
    if (threadIdx.x < 64 && c < H_pad)
      part[(int64_t)blockIdx.x * H_pad + c] = (psum[0][threadIdx.x] + psum[1][threadIdx.x]) + (psum[2][threadIdx.x] + psum[3][threadIdx.x]);
  }
}

extern "C" int64_t xdfm_cin_dy_db_workspace_bytes(int64_t B, int D, int H_pad) { return ceil_div64(B * (int64_t)D, 64) * H_pad * 4; }

extern "C" int xdfm_cin_dy_rows_cols_db(const void* yt, int64_t B, int D, int H, int Hs, int H_pad, int direct_begin, const float* dpooled,
                                        const float* dmaps, int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next,
                                        int act, void* dyt, void* dyT, float* db, void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int xdfm_cin_dy_rows_cols(const void* yt, int64_t B, int D, int H, int Hs, int H_pad, int direct_begin, const float* dpooled,
                                     const float* dmaps, int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next,
                                     int act, void* dyt, void* dyT, void* stream) {
  return xdfm_cin_dy_rows_cols_db(yt, B, D, H, Hs, H_pad, direct_begin, dpooled, dmaps, fm_total, col_off, dnext, dnext_pitch, n_next, act,
                                  dyt, dyT, nullptr, nullptr, 0, stream);
}

// As xdfm_cin_dy_rows_cols; db != NULL also produces the layer's bias gradient db[h] = sum_r dY[r, h] (of the bf16 values, fixed
// order) from per-tile partial sums in `workspace` (xdfm_cin_dy_db_workspace_bytes) -- the separate pass over dyT is gone.
extern "C" int xdfm_cin_dy_rows_cols_db(const void* yt, int64_t B, int D, int H, int Hs, int H_pad, int direct_begin, const float* dpooled,
                                     const float* dmaps, int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next,
                                        int act, void* dyt, void* dyT, float* db, void* workspace, int64_t workspace_bytes, void* stream) {
  XDFM_CHECK_ARG(Hs % 8 == 0 && Hs >= H && H_pad >= Hs, "cin_dy_rows_cols: Hs=%d (multiple of 8, >= H=%d), H_pad=%d >= Hs", Hs, H, H_pad);
  XDFM_CHECK_ARG(db == nullptr || (workspace != nullptr && workspace_bytes >= xdfm_cin_dy_db_workspace_bytes(B, D, H_pad)),
                 "cin_dy_rows_cols: workspace too small for the bias gradient");
  XDFM_CHECK_ARG(act == XDFM_ACT_RELU || act == XDFM_ACT_NONE, "cin_dy_rows_cols: activation %d not supported on the bf16 path", act);
  const int64_t R = B * (int64_t)D;
  if (R == 0) return XDFM_OK;
  XDFM_CHECK_ARG(R < ((int64_t)1 << 32), "cin_dy_rows_cols: B * D = %lld rows exceed the 32-bit row arithmetic of the kernel", (long long)R);
  dim3 grid((unsigned)ceil_div64(R, 64), (unsigned)ceil_div64(H_pad, 64));
  cin_dy_rows_cols_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)yt, R, D, H, Hs, H_pad, direct_begin, dpooled, dmaps,
                                                                  fm_total, col_off, dnext, dnext_pitch, n_next, act, (__nv_bfloat16*)dyt,
                                                                  (__nv_bfloat16*)dyT, db != nullptr ? (float*)workspace : nullptr);
  XDFM_LAUNCH_CHECK();
  if (db != nullptr) {
    XDFM_TILE_COLSUM((const float*)workspace, ceil_div64(R, 64), (int64_t)H_pad, H, db, 0, (cudaStream_t)stream);   // fixed order
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}
