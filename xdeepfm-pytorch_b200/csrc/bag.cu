// Multi-value (VarLenSparseFeat) pooling: the bag mode of the embedding lookup.
//
// Replaces (reference, file:line):
//   varlen_embedding_lookup                        deepctr/inputs.py:212-225     (one nn.Embedding call per sequence feature)
//   get_varlen_pooling_list                        deepctr/inputs.py:141-155     (mask = ids != 0, or positions < length column)
//   SequencePoolingLayer.forward (sum/mean/max)    deepctr/layers/sequence.py:51-79
//   cat(sparse_embedding_list + varlen list)       deepctr/models/basemodel.py:354-380, xdeepfm.py:86
// The reference materialises, per sequence feature, the [B, T, E] gather, a repeat_interleave'd [B, T, E] mask, the masked
// product and the reduction (4 HBM round trips of the sequence tensor).  Here the T positions of a sequence feature are T
// "slots" of the ONE fused multi-table gather (embed.cu), and this kernel turns the [B, S, D] slot tensor into the [B, F, D]
// field tensor the CIN / DNN read: fixed fields are copied, sequence fields are reduced under their mask -- one read of the
// slot tensor, one write of the field tensor.  HBM-bound: (S + F) * D * 4 bytes per sample forward, the same backward.
#include "common.cuh"
#include "../../include/xdfm.h"

#define BAG_MAX 64

struct BagLayout {
  int32_t slot0[BAG_MAX];    // first slot of field f
  int32_t slen[BAG_MAX];     // number of slots of field f (1 for a fixed field)
  int32_t mode[BAG_MAX];     // XDFM_BAG_SINGLE / SUM / MEAN / MAX
  int32_t lencol[BAG_MAX];   // column of `lens` holding the sequence length of field f, or -1: mask = (id != 0)
  int8_t field_of[BAG_MAX];  // slot -> field
};


// 128-bit accesses when D % 4 == 0 (rows of the slot / field tensors are then 16-byte aligned), scalar otherwise (D = 1: first-order term)
template <int VEC, typename T>
__device__ __forceinline__ void ldv(const T* __restrict__ p, T (&x)[VEC]) {
  if constexpr (VEC == 4) {
    const int4 t = *reinterpret_cast<const int4*>(p);
    x[0] = *reinterpret_cast<const T*>(&t.x); x[1] = *reinterpret_cast<const T*>(&t.y);
    x[2] = *reinterpret_cast<const T*>(&t.z); x[3] = *reinterpret_cast<const T*>(&t.w);
  } else {
    x[0] = *p;
  }
}
template <int VEC, typename T>
__device__ __forceinline__ void stv(T* __restrict__ p, const T (&x)[VEC]) {
  if constexpr (VEC == 4) {
    int4 t;
    t.x = *reinterpret_cast<const int*>(&x[0]); t.y = *reinterpret_cast<const int*>(&x[1]);
    t.z = *reinterpret_cast<const int*>(&x[2]); t.w = *reinterpret_cast<const int*>(&x[3]);
    *reinterpret_cast<int4*>(p) = t;
  } else {
    *p = x[0];
  }
}

// validity bitmask of the positions of field f for one sample (bit j = position j counts) and the divisor of 'mean'
// (sequence.py:53-61): supports_masking: mask = ids != 0, length = sum(mask);  else mask = arange(T) < length, length = the length
// column as given.  All id loads are independent (no load -> branch -> load chain); S <= 64 so the mask fits one register pair.
__device__ __forceinline__ uint64_t bag_mask(const BagLayout& lay, int f, const int32_t* __restrict__ idrow,
                                             const int32_t* __restrict__ lenrow, float* count) {
  const int lc = lay.lencol[f], L = lay.slen[f];
  if (lc >= 0) {
    const int len = lenrow[lc];
    *count = (float)len;
    const int n = min(max(len, 0), L);
    return n >= 64 ? ~0ull : ((1ull << n) - 1ull);
  }
  const int32_t* p = idrow + lay.slot0[f];
  uint64_t m = 0;
#pragma unroll 4
  for (int j = 0; j < L; ++j) m |= (uint64_t)(p[j] != 0) << j;
  *count = (float)__popcll(m);
  return m;
}

// IT = index type of the flattened (sample, field / slot, vector) loop: uint32_t whenever the launch has < 2^31 items (64-bit
// divisions cost more than the memory accesses of a thread), int64_t otherwise
template <int VEC, typename IT>
__global__ void __launch_bounds__(256) bag_pool_fwd_kernel(const float* __restrict__ emb, const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ lens, int nlen, int64_t B, int S, int D, int F,
                                                           const __grid_constant__ BagLayout lay, float* __restrict__ out,
                                                           int32_t* __restrict__ argmax, float* __restrict__ den) {
  const int DV = D / VEC;
  const IT total = (IT)(B * F * DV), per_sample = (IT)(DV * F);
  for (IT t = (IT)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (IT)gridDim.x * blockDim.x) {
    const int64_t b = (int64_t)(t / per_sample);
    const int rem = (int)(t - (IT)b * per_sample);
    const int f = rem / DV, dv = rem - f * DV;
    const int s0 = lay.slot0[f], L = lay.slen[f], mode = lay.mode[f];
    const float* src = emb + ((b * S + s0) * D + dv * VEC);
    const int32_t* idrow = ids + b * S;
    const int32_t* lenrow = lens ? lens + b * nlen : nullptr;
    float acc[VEC];
    int32_t arg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { acc[v] = 0.f; arg[v] = 0; }
    if (mode == XDFM_BAG_SINGLE) {
      ldv<VEC>(src, acc);
    } else {
      float cnt;
      const uint64_t mask = bag_mask(lay, f, idrow, lenrow, &cnt);
      if (mode == XDFM_BAG_MAX) {
        // hist = x - (1 - mask) * 1e9; max over the positions (sequence.py:69-72); first position wins ties.  Masked rows are
        // read too: fl(x - 1e9) is part of the reference's result
#pragma unroll 4
        for (int j = 0; j < L; ++j) {
          const float pen = ((mask >> j) & 1ull) ? 0.f : 1e9f;
          float xs[VEC];
          ldv<VEC>(src + (int64_t)j * D, xs);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float x = xs[v] - pen;
            if (j == 0 || x > acc[v]) { acc[v] = x; arg[v] = j; }
          }
        }
      } else {
        // rows of padded positions are not read (predicated loads, four in flight)
#pragma unroll 4
        for (int j = 0; j < L; ++j) {
          float xs[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) xs[v] = 0.f;
          if ((mask >> j) & 1ull) ldv<VEC>(src + (int64_t)j * D, xs);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += xs[v];
        }
        if (mode == XDFM_BAG_MEAN) {
          const float dn = cnt + 1e-8f;      // sequence.py:76-77
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = acc[v] / dn;
          if (dv == 0) den[b * F + f] = dn;
        }
      }
    }
    stv<VEC>(out + ((b * F + f) * D + dv * VEC), acc);
    if (argmax != nullptr) stv<VEC>(argmax + ((b * F + f) * D + dv * VEC), arg);
  }
}

template <int VEC, typename IT>
__global__ void __launch_bounds__(256) bag_pool_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ lens, int nlen,
                                                           const int32_t* __restrict__ argmax, const float* __restrict__ den, int64_t B,
                                                           int S, int D, int F, const __grid_constant__ BagLayout lay,
                                                           float* __restrict__ demb) {
  const int DV = D / VEC;
  const IT total = (IT)(B * S * DV), per_sample = (IT)(DV * S);
  for (IT t = (IT)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (IT)gridDim.x * blockDim.x) {
    const int64_t b = (int64_t)(t / per_sample);
    const int rem = (int)(t - (IT)b * per_sample);
    const int s = rem / DV, dv = rem - s * DV;
    const int f = lay.field_of[s];
    const int j = s - lay.slot0[f], mode = lay.mode[f];
    const int32_t* idrow = ids + b * S;
    const int32_t* lenrow = lens ? lens + b * nlen : nullptr;
    float g[VEC], r[VEC];
    ldv<VEC>(dout + ((b * F + f) * D + dv * VEC), g);
#pragma unroll
    for (int v = 0; v < VEC; ++v) r[v] = 0.f;
    if (mode == XDFM_BAG_SINGLE) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = g[v];
    } else if (mode == XDFM_BAG_MAX) {
      int32_t am[VEC];
      ldv<VEC>(argmax + ((b * F + f) * D + dv * VEC), am);
#pragma unroll
      for (int v = 0; v < VEC; ++v) r[v] = (am[v] == j) ? g[v] : 0.f;
    } else {
      const int lc = lay.lencol[f];
      const bool valid = lc >= 0 ? (j < lenrow[lc]) : (idrow[s] != 0);
      if (valid) {
        if (mode == XDFM_BAG_MEAN) {
          const float dn = den[b * F + f];        // count + 1e-8, written by the forward
#pragma unroll
          for (int v = 0; v < VEC; ++v) r[v] = g[v] / dn;
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) r[v] = g[v];
        }
      }
    }
    stv<VEC>(demb + ((b * S + s) * D + dv * VEC), r);
  }
}

static int bag_layout(const int32_t* slot0, const int32_t* slen, const int32_t* mode, const int32_t* lencol, int F, int S, int nlen,
                      bool has_lens, const void* argmax, const void* den, BagLayout* lay) {
  XDFM_CHECK_ARG(F >= 1 && F <= BAG_MAX && S >= F && S <= BAG_MAX, "bag_pool: need 1 <= F <= S <= %d (F %d, S %d)", BAG_MAX, F, S);
  XDFM_CHECK_ARG(slot0 && slen && mode && lencol, "bag_pool: NULL layout array");
  int next = 0;
  bool any_max = false, any_mean = false;
  for (int f = 0; f < F; ++f) {
    XDFM_CHECK_ARG(slot0[f] == next && slen[f] >= 1, "bag_pool: field %d: slots must be contiguous and non-empty (slot0 %d, len %d)", f,
                   slot0[f], slen[f]);
    XDFM_CHECK_ARG(mode[f] >= XDFM_BAG_SINGLE && mode[f] <= XDFM_BAG_MAX, "bag_pool: field %d: unknown mode %d", f, mode[f]);
    XDFM_CHECK_ARG(mode[f] != XDFM_BAG_SINGLE || slen[f] == 1, "bag_pool: field %d: a fixed field has exactly one slot", f);
    XDFM_CHECK_ARG(lencol[f] < nlen && (lencol[f] < 0 || has_lens), "bag_pool: field %d: length column %d of %d", f, lencol[f], nlen);
    XDFM_CHECK_ARG(next + slen[f] <= BAG_MAX, "bag_pool: more than %d slots", BAG_MAX);
    lay->slot0[f] = slot0[f];
    lay->slen[f] = slen[f];
    lay->mode[f] = mode[f];
    lay->lencol[f] = lencol[f];
    for (int j = 0; j < slen[f]; ++j) lay->field_of[next + j] = (int8_t)f;
    next += slen[f];
    any_max |= mode[f] == XDFM_BAG_MAX;
    any_mean |= mode[f] == XDFM_BAG_MEAN;
  }
  XDFM_CHECK_ARG(next == S, "bag_pool: the fields cover %d slots, the slot tensor has %d", next, S);
  XDFM_CHECK_ARG(!any_max || argmax != nullptr, "bag_pool: a 'max' field needs the argmax buffer");
  XDFM_CHECK_ARG(!any_mean || den != nullptr, "bag_pool: a 'mean' field needs the divisor buffer");
  return XDFM_OK;
}

static inline bool bag_al16(const void* p) { return ((uintptr_t)p & 15) == 0; }      // NULL counts as aligned

static inline int bag_grid(int64_t total) {
  const int64_t want = ceil_div64(total, 256);
  const int64_t cap = (int64_t)xdfm_num_sms() * 64;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

extern "C" int xdfm_bag_pool_fwd(const float* emb, const int32_t* ids, const int32_t* lens, int nlen, int64_t B, int S, int D, int F,
                                 const int32_t* slot0, const int32_t* slen, const int32_t* mode, const int32_t* lencol, float* out,
                                 int32_t* argmax, float* den, void* stream) {
  XDFM_CHECK_ARG(B >= 0 && D >= 1 && nlen >= 0, "bag_pool_fwd: bad sizes (B %lld, D %d, nlen %d)", (long long)B, D, nlen);
  BagLayout lay;
  int rc = bag_layout(slot0, slen, mode, lencol, F, S, nlen, lens != nullptr, argmax, den, &lay);
  if (rc != XDFM_OK) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CHECK_ARG(emb && ids && out, "bag_pool_fwd: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (D % 4 == 0 && bag_al16(emb) && bag_al16(out) && bag_al16(argmax)) ? 4 : 1;     // views may start mid-row
  const int64_t total = B * F * (D / vec);
  const bool small = total < (1ll << 31);
#define BAG_FWD(V, IT) bag_pool_fwd_kernel<V, IT><<<bag_grid(total), 256, 0, st>>>(emb, ids, lens, nlen, B, S, D, F, lay, out, argmax, den)
  if (vec == 4) { if (small) BAG_FWD(4, uint32_t); else BAG_FWD(4, int64_t); }
  else { if (small) BAG_FWD(1, uint32_t); else BAG_FWD(1, int64_t); }
#undef BAG_FWD
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_bag_pool_bwd(const float* dout, const int32_t* ids, const int32_t* lens, int nlen, const int32_t* argmax,
                                 const float* den, int64_t B, int S, int D, int F, const int32_t* slot0, const int32_t* slen, const int32_t* mode,
                                 const int32_t* lencol, float* demb, void* stream) {
  XDFM_CHECK_ARG(B >= 0 && D >= 1 && nlen >= 0, "bag_pool_bwd: bad sizes (B %lld, D %d, nlen %d)", (long long)B, D, nlen);
  BagLayout lay;
  int rc = bag_layout(slot0, slen, mode, lencol, F, S, nlen, lens != nullptr, argmax, den, &lay);
  if (rc != XDFM_OK) return rc;
  if (B == 0) return XDFM_OK;
  XDFM_CHECK_ARG(dout && ids && demb, "bag_pool_bwd: NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = (D % 4 == 0 && bag_al16(dout) && bag_al16(demb) && bag_al16(argmax)) ? 4 : 1;
  const int64_t total = B * S * (D / vec);
  const bool small = total < (1ll << 31);
#define BAG_BWD(V, IT) bag_pool_bwd_kernel<V, IT><<<bag_grid(total), 256, 0, st>>>(dout, ids, lens, nlen, argmax, den, B, S, D, F, lay, demb)
  if (vec == 4) { if (small) BAG_BWD(4, uint32_t); else BAG_BWD(4, int64_t); }
  else { if (small) BAG_BWD(1, uint32_t); else BAG_BWD(1, int64_t); }
#undef BAG_BWD
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
