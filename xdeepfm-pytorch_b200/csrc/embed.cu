// Sparse embedding path: input split, fused multi-table gather (+ first-order linear term),
// deterministic sorted segmented scatter-add of the row gradients.
//
// Replaces (reference, file:line):
//   X[:, i:i+1].long() + 26x nn.Embedding lookups + cat      deepctr/models/basemodel.py:354-380, xdeepfm.py:86
//   Linear.forward (26x [V,1] lookups, sum, dense @ weight)  deepctr/models/basemodel.py:63-92
//   embedding_dense_backward (dense [V,D] grad per table)    torch autograd of the above
//
// HBM-bound integer/byte work: coalesced 128-bit row reads/writes, grid sized in multiples of the SM
// count, no tensor cores.  Algorithmic bytes per looked-up row: 4 (id) + D*4 (row read) + D*4 (write).
#include "common.cuh"
#include <cub/cub.cuh>
#include "../../include/xdfm.h"

struct FieldPtrs {
  const float* p[XDFM_MAX_FIELDS];
  const float* lin[XDFM_MAX_FIELDS];
  int32_t vocab[XDFM_MAX_FIELDS];
};

struct ColMap {
  int32_t sparse[XDFM_MAX_FIELDS];
  int32_t dense[XDFM_MAX_DENSE];
};

// ------------------------------------------------------------------------------------------------
// split_input: ids = (int)X[:, sparse_col] (truncation == .long()), dense = X[:, dense_col]
// ------------------------------------------------------------------------------------------------
__global__ void split_input_kernel(const float* __restrict__ X, int64_t B, int ncol, ColMap cm, int m, int nd,
                                   int32_t* __restrict__ ids, float* __restrict__ dense) {
  int64_t total = B * (int64_t)(m + nd);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / (m + nd);
    int c = (int)(i - b * (m + nd));
    if (c < m) {
      ids[b * m + c] = (int32_t)X[b * ncol + cm.sparse[c]];
    } else {
      dense[b * nd + (c - m)] = X[b * ncol + cm.dense[c - m]];
    }
  }
}

extern "C" int xdfm_split_input(const float* X, int64_t B, int ncol, const int32_t* sparse_cols, int m,
                                const int32_t* dense_cols, int nd, int32_t* ids, float* dense, void* stream) {
  XDFM_CHECK_ARG(m >= 0 && m <= XDFM_MAX_FIELDS, "split_input: m=%d out of range (max %d)", m, XDFM_MAX_FIELDS);
  XDFM_CHECK_ARG(nd >= 0 && nd <= XDFM_MAX_DENSE, "split_input: nd=%d out of range (max %d)", nd, XDFM_MAX_DENSE);
  if (B == 0 || m + nd == 0) return XDFM_OK;
  ColMap cm;
  for (int i = 0; i < m; ++i) {
    XDFM_CHECK_ARG(sparse_cols[i] >= 0 && sparse_cols[i] < ncol, "split_input: sparse col %d out of range", sparse_cols[i]);
    cm.sparse[i] = sparse_cols[i];
  }
  for (int i = 0; i < nd; ++i) {
    XDFM_CHECK_ARG(dense_cols[i] >= 0 && dense_cols[i] < ncol, "split_input: dense col %d out of range", dense_cols[i]);
    cm.dense[i] = dense_cols[i];
  }
  int64_t total = B * (int64_t)(m + nd);
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  split_input_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, B, ncol, cm, m, nd, ids, dense);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// fused multi-table gather: out[b, f, :] = table_f[ids[b, f], :]
// One thread moves one VEC-wide piece of a row; 4 independent rows in flight per thread.
// ------------------------------------------------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<4> { typedef float4 T; };
template <>
struct VecT<2> { typedef float2 T; };
template <>
struct VecT<1> { typedef float T; };

template <int VEC>
__global__ void __launch_bounds__(256) embed_gather_kernel(FieldPtrs fp, const int32_t* __restrict__ ids, int64_t n_rows,
                                                           int m, int D, float* __restrict__ out) {
  typedef typename VecT<VEC>::T V;
  const int vpr = D / VEC;  // vectors per row
  const int64_t total = n_rows * vpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    V val[U];
    int64_t dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      dst[u] = -1;
      if (i < total) {
        int64_t row = i / vpr;
        int v = (int)(i - row * vpr);
        int f = (int)(row % m);
        int id = __ldg(ids + row);
        id = max(0, min(id, fp.vocab[f] - 1));
        const V* src = reinterpret_cast<const V*>(fp.p[f] + (int64_t)id * D) + v;
        val[u] = __ldg(src);
        dst[u] = i;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (dst[u] >= 0) reinterpret_cast<V*>(out)[dst[u]] = val[u];
    }
  }
}

// first-order term: lin[b] = sum_f lin_f[ids[b,f]] + sum_j dense[b,j]*w[j]; one warp per sample, shuffle reduce
__global__ void __launch_bounds__(256) linear_term_kernel(FieldPtrs fp, const int32_t* __restrict__ ids, int64_t B, int m,
                                                          const float* __restrict__ dense, int nd,
                                                          const float* __restrict__ dense_w, float* __restrict__ out_lin) {
  int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float acc = 0.f;
    for (int f = lane; f < m; f += 32) {
      int id = __ldg(ids + b * m + f);
      id = max(0, min(id, fp.vocab[f] - 1));
      acc += __ldg(fp.lin[f] + id);
    }
    float accd = 0.f;
    if (dense_w != nullptr) {
      for (int j = lane; j < nd; j += 32) accd += __ldg(dense + b * nd + j) * __ldg(dense_w + j);
    }
    // the reference adds the sparse sum first, then the dense matmul (basemodel.py:81-90)
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) out_lin[b] = acc + accd;
  }
}

extern "C" int xdfm_embed_gather(const float* const* tables, const float* const* lin_tables, const int32_t* vocab,
                                 const int32_t* ids, int64_t B, int m, int D, float* out_emb, const float* dense, int nd,
                                 const float* dense_w, float* out_lin, void* stream) {
  XDFM_CHECK_ARG(m >= 0 && m <= XDFM_MAX_FIELDS, "embed_gather: m=%d out of range (max %d)", m, XDFM_MAX_FIELDS);
  XDFM_CHECK_ARG(D >= 1, "embed_gather: D=%d", D);
  if (B == 0) return XDFM_OK;
  FieldPtrs fp;
  for (int f = 0; f < m; ++f) {
    fp.p[f] = tables ? tables[f] : nullptr;
    fp.lin[f] = lin_tables ? lin_tables[f] : nullptr;
    fp.vocab[f] = vocab[f];
    XDFM_CHECK_ARG(vocab[f] > 0, "embed_gather: vocab[%d]=%d", f, vocab[f]);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (out_emb != nullptr && m > 0) {
    XDFM_CHECK_ARG(tables != nullptr, "embed_gather: tables is null");
    int64_t n_rows = B * (int64_t)m;
    bool al16 = (D % 4 == 0) && ((uintptr_t)out_emb % 16 == 0);
    bool al8 = (D % 2 == 0) && ((uintptr_t)out_emb % 8 == 0);
    for (int f = 0; f < m; ++f) {
      al16 = al16 && ((uintptr_t)fp.p[f] % 16 == 0);
      al8 = al8 && ((uintptr_t)fp.p[f] % 8 == 0);
    }
    int vec = al16 ? 4 : (al8 ? 2 : 1);
    int64_t total = n_rows * (D / vec);
    int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(total, 256 * 4));
    blocks = max(blocks, 1);
    if (vec == 4) embed_gather_kernel<4><<<blocks, 256, 0, st>>>(fp, ids, n_rows, m, D, out_emb);
    else if (vec == 2) embed_gather_kernel<2><<<blocks, 256, 0, st>>>(fp, ids, n_rows, m, D, out_emb);
    else embed_gather_kernel<1><<<blocks, 256, 0, st>>>(fp, ids, n_rows, m, D, out_emb);
    XDFM_LAUNCH_CHECK();
  }
  if (out_lin != nullptr) {
    XDFM_CHECK_ARG(m == 0 || lin_tables != nullptr, "embed_gather: lin_tables is null");
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(B, 8));
    linear_term_kernel<<<max(blocks, 1), 256, 0, st>>>(fp, ids, B, m, dense, nd, (nd > 0 ? dense_w : nullptr), out_lin);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// backward: sort (table-row keys) -> run-length segments -> fixed-order segmented reduce
// ------------------------------------------------------------------------------------------------
struct RowOffsets {
  int64_t off[XDFM_MAX_FIELDS];
  int32_t vocab[XDFM_MAX_FIELDS];
};

__global__ void make_keys_kernel(const int32_t* __restrict__ ids, int64_t n, int m, RowOffsets ro, uint32_t* __restrict__ keys,
                                 int32_t* __restrict__ pos) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int f = (int)(i % m);
    int id = max(0, min(ids[i], ro.vocab[f] - 1));
    keys[i] = (uint32_t)(ro.off[f] + id);
    pos[i] = (int32_t)i;
  }
}

// seg_offsets[0..num] = exclusive prefix of run lengths; also writes seg_offsets[num] = n
__global__ void finalize_offsets_kernel(int32_t* seg_offsets, const int32_t* num_segments, int32_t n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) seg_offsets[*num_segments] = n;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BwdWsLayout {
  size_t keys_in, pos_in, counts, cub_temp, cub_bytes, total;
};

static BwdWsLayout bwd_ws_layout(int64_t n) {
  BwdWsLayout L;
  size_t o = 0;
  L.keys_in = o; o += align256(n * sizeof(uint32_t));
  L.pos_in = o; o += align256(n * sizeof(int32_t));
  L.counts = o; o += align256((n + 1) * sizeof(int32_t));
  size_t sort_b = 0, rle_b = 0, scan_b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, 32);
  cub::DeviceRunLengthEncode::Encode(nullptr, rle_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                     (int32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  L.cub_bytes = align256(max(sort_b, max(rle_b, scan_b)));
  L.cub_temp = o; o += L.cub_bytes;
  L.total = o;
  return L;
}

extern "C" int64_t xdfm_embed_bwd_workspace_bytes(int64_t n_keys) {
  if (n_keys <= 0) return 256;
  return (int64_t)bwd_ws_layout(n_keys).total;
}

extern "C" int xdfm_embed_bwd_segments(const int32_t* ids, int64_t B, int m, const int64_t* row_offset, const int32_t* vocab,
                                       int64_t total_rows, void* workspace, int64_t workspace_bytes, uint32_t* uniq_keys,
                                       int32_t* seg_offsets, int32_t* sorted_pos, int32_t* num_segments, void* stream) {
  XDFM_CHECK_ARG(m >= 1 && m <= XDFM_MAX_FIELDS, "embed_bwd_segments: m=%d", m);
  int64_t n = B * (int64_t)m;
  XDFM_CHECK_ARG(n < (int64_t)1 << 31, "embed_bwd_segments: B*m too large");
  XDFM_CHECK_ARG(total_rows < ((int64_t)1 << 32), "embed_bwd_segments: total_rows >= 2^32 needs sharding");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    XDFM_CUDA(cudaMemsetAsync(num_segments, 0, sizeof(int32_t), st));
    return XDFM_OK;
  }
  BwdWsLayout L = bwd_ws_layout(n);
  XDFM_CHECK_ARG(workspace_bytes >= (int64_t)L.total, "embed_bwd_segments: workspace too small (%lld < %lld)",
                 (long long)workspace_bytes, (long long)L.total);
  char* ws = (char*)workspace;
  uint32_t* keys_in = (uint32_t*)(ws + L.keys_in);
  int32_t* pos_in = (int32_t*)(ws + L.pos_in);
  int32_t* counts = (int32_t*)(ws + L.counts);
  void* cub_temp = ws + L.cub_temp;
  RowOffsets ro;
  for (int f = 0; f < m; ++f) { ro.off[f] = row_offset[f]; ro.vocab[f] = vocab[f]; }
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  make_keys_kernel<<<blocks, 256, 0, st>>>(ids, n, m, ro, keys_in, pos_in);
  XDFM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((int64_t)1 << end_bit) < total_rows) ++end_bit;
  // sort keys into uniq_keys (used as scratch), copy them back over keys_in, then run-length encode keys_in -> uniq_keys
  size_t tb = L.cub_bytes;
  XDFM_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, tb, (const uint32_t*)keys_in, uniq_keys, (const int32_t*)pos_in,
                                            sorted_pos, (int)n, 0, end_bit, st));
  // keys_in is free now: copy sorted keys there and RLE from it into uniq_keys
  XDFM_CUDA(cudaMemcpyAsync(keys_in, uniq_keys, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  tb = L.cub_bytes;
  XDFM_CUDA(cub::DeviceRunLengthEncode::Encode(cub_temp, tb, (const uint32_t*)keys_in, uniq_keys, counts, num_segments, (int)n, st));
  tb = L.cub_bytes;
  // exclusive scan over all n slots (entries past num_segments are garbage but never read)
  XDFM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp, tb, (const int32_t*)counts, seg_offsets, (int)n, st));
  finalize_offsets_kernel<<<1, 32, 0, st>>>(seg_offsets, num_segments, (int32_t)n);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// Fixed-order segmented reduce.  A "slot group" of LPR lanes (LPR = D/VEC, power of two <= 32) owns one row
// of D floats; a warp has S = 32/LPR slot groups.  Segment entries e = 0..len-1 are dealt round-robin to
// the S slots of ONE warp (short segments) or of ALL warps of a block (long segments); every slot adds its
// entries in increasing e, then slots are combined by a fixed xor-shuffle tree (and a fixed smem order
// across warps).  The summation order depends only on (len, D), never on scheduling -> bit-reproducible.
#define SEG_LONG 256
#define SEG_TINY 16

template <int VEC>
__device__ __forceinline__ void vec_add(float* a, const float* b) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] += b[i];
}

// One slot's share of a segment: entries first, first + stride, ... < end, added in increasing order.  Four independent
// (position -> row) load chains are in flight per iteration; the additions still happen in entry order, so the result is the
// same fixed function of (segment length, D) as a plain sequential loop.
template <int VEC>
__device__ __forceinline__ void seg_accumulate(const float* __restrict__ demb, const float* __restrict__ dlin,
                                               const int32_t* __restrict__ sorted_pos, int first, int end, int stride, int m, int D,
                                               int sub, bool active, float (&acc)[VEC], float& accl) {
  typedef typename VecT<VEC>::T V;
  constexpr int U = 4;
  int e = first;
  for (; e + (U - 1) * stride < end; e += U * stride) {
    int p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = __ldg(sorted_pos + e + u * stride);
    V v[U];
    float l[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (active && demb != nullptr) v[u] = __ldg(reinterpret_cast<const V*>(demb + (int64_t)p[u] * D) + sub);
      l[u] = (sub == 0 && dlin != nullptr) ? __ldg(dlin + p[u] / m) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (active && demb != nullptr) vec_add<VEC>(acc, reinterpret_cast<const float*>(&v[u]));
      accl += l[u];
    }
  }
  for (; e < end; e += stride) {
    const int p = __ldg(sorted_pos + e);
    if (active && demb != nullptr) {
      V v = __ldg(reinterpret_cast<const V*>(demb + (int64_t)p * D) + sub);
      vec_add<VEC>(acc, reinterpret_cast<const float*>(&v));
    }
    if (sub == 0 && dlin != nullptr) accl += __ldg(dlin + p / m);
  }
}

// Long segments (hot rows: a field with 3 distinct ids has ~B/3 entries per row) are found by the short pass and cut into CHUNKS
// of SEG_CHUNK entries that are appended to a per-device work list; the long pass reduces one chunk per block into a partial row,
// and the block that finishes a segment's last outstanding chunk adds the partials in chunk order.  One block per WHOLE segment
// (round 1) left the step waiting for the longest segment: at B = 65 536, D = 64 a 40 000-entry segment took 1.7 ms on one block
// while the other 147 SMs idled (4 % of the HBM peak for the whole reduce).  The list order depends on scheduling, the sums do
// not: every chunk is reduced in a fixed order, and the chunks of a segment are combined in increasing chunk index.
#define SEG_CHUNK 512
#define SEG_ITEMS_MAX 16384
#define SEG_PART_FLOATS (1 << 20)
__device__ int g_seg_item_count;
__device__ int4 g_seg_items[SEG_ITEMS_MAX];          // (segment, chunk, chunks of the segment, first item of the segment)
__device__ int g_seg_done[SEG_ITEMS_MAX];            // per segment (indexed by its first item): chunks finished
__device__ float g_seg_part[SEG_PART_FLOATS];        // [item][D] partial rows
__device__ float g_seg_part_lin[SEG_ITEMS_MAX];
// launches whose chunk count could exceed the static work list fall back to one block per whole segment, found by scanning

// long_pass: 0 short segments + chunk list of the long ones | 3 reduce the listed chunks | legacy: -1 short only (no list),
// 2 one block per long segment found by scanning every segment
template <int VEC>
__global__ void __launch_bounds__(256, 4) seg_reduce_kernel(const float* __restrict__ demb, const float* __restrict__ dlin,
                                                         const int32_t* __restrict__ sorted_pos,
                                                         const int32_t* __restrict__ seg_offsets,
                                                         const int32_t* __restrict__ num_segments, int m, int D, int lpr,
                                                         float* __restrict__ gsum, float* __restrict__ gsum_lin, int long_pass) {
  typedef typename VecT<VEC>::T V;
  const int nseg = *num_segments;
  const int lane = threadIdx.x & 31;
  const int S = 32 / lpr;
  const int slot = lane / lpr;
  const int sub = lane % lpr;           // which VEC piece of the row (valid if sub*VEC < D)
  const bool active = sub * VEC < D;
  __shared__ float sh[8][32 * 4 + 8];
  __shared__ int sh_last;
  if (long_pass <= 0) {
    // S segments per warp at a time.  TINY segments (<= SEG_TINY entries: nearly all of them in a Criteo-shaped batch, where the
    // average row is looked up 4 times) are summed by ONE slot group each -- S independent (offsets -> positions -> rows) chains per
    // warp instead of one, up to four rows in flight per group -- in entry order.  The others take the whole warp, one after another
    // (entries dealt round-robin to the S slots, fixed shuffle tree), or go to the long pass.
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // (a warp's S segments are nwarps apart: sorted keys put the medium-length segments of a low-cardinality table next to each
    // other, and a warp that drew eight of them in a row worked through them alone while the others had finished)
    for (int64_t s0 = warp; s0 < nseg; s0 += nwarps * S) {
      const int64_t sg = s0 + (int64_t)slot * nwarps;
      int beg = 0, end = 0;
      if (sg < nseg) {
        beg = seg_offsets[sg];
        end = seg_offsets[sg + 1];
      }
      const int len = end - beg;
      if (len > 0 && len <= SEG_TINY) {
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        float accl = 0.f;
        for (int e = beg; e < end; e += 4) {
          int p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) p[u] = e + u < end ? __ldg(sorted_pos + e + u) : -1;
          V v[4];
          float l[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            l[u] = 0.f;
            if (p[u] >= 0) {
              if (active && demb != nullptr) v[u] = __ldg(reinterpret_cast<const V*>(demb + (int64_t)p[u] * D) + sub);
              if (sub == 0 && dlin != nullptr) l[u] = __ldg(dlin + p[u] / m);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (p[u] >= 0) {
              if (active && demb != nullptr) vec_add<VEC>(acc, reinterpret_cast<const float*>(&v[u]));
              accl += l[u];
            }
          }
        }
        if (active && gsum != nullptr) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) gsum[sg * (int64_t)D + sub * VEC + i] = acc[i];
        }
        if (sub == 0 && gsum_lin != nullptr) gsum_lin[sg] = accl;
      }
      unsigned pending = __ballot_sync(0xffffffffu, len > SEG_TINY && sub == 0);
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int64_t s = s0 + (int64_t)(src / lpr) * nwarps;
        const int wb = __shfl_sync(0xffffffffu, beg, src), we = __shfl_sync(0xffffffffu, end, src);
        if (we - wb > SEG_LONG) {          // handled by the long pass
          if (long_pass == 0) {
            const int nch = (we - wb + SEG_CHUNK - 1) / SEG_CHUNK;
            int base = 0;
            if (lane == 0) base = atomicAdd(&g_seg_item_count, nch);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int c = lane; c < nch; c += 32) g_seg_items[base + c] = make_int4((int)s, c, nch, base);
            if (lane == 0) g_seg_done[base] = 0;
          }
          continue;
        }
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        float accl = 0.f;
        seg_accumulate<VEC>(demb, dlin, sorted_pos, wb + slot, we, S, m, D, sub, active, acc, accl);
        for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
          accl += __shfl_xor_sync(0xffffffffu, accl, o);
        }
        if (slot == 0) {
          if (active && gsum != nullptr) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) gsum[s * (int64_t)D + sub * VEC + i] = acc[i];
          }
          if (sub == 0 && gsum_lin != nullptr) gsum_lin[s] = accl;
        }
      }
    }
  } else {
    // long segments: one block per chunk (mode 3) or per whole segment (mode 2), grid-stride; 8 warps x S slots
    const int w = threadIdx.x >> 5;
    const bool chunked = long_pass == 3;
    const int64_t n_iter = chunked ? g_seg_item_count : nseg;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
      int4 item = chunked ? g_seg_items[it] : make_int4((int)it, 0, 1, 0);
      const int64_t s = item.x;
      int beg = seg_offsets[s], end = seg_offsets[s + 1];
      if (end - beg <= SEG_LONG) continue;
      if (chunked) {
        beg += item.y * SEG_CHUNK;
        end = min(end, beg + SEG_CHUNK);
      }
      float acc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
      float accl = 0.f;
      seg_accumulate<VEC>(demb, dlin, sorted_pos, beg + w * S + slot, end, 8 * S, m, D, sub, active, acc, accl);
      for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        accl += __shfl_xor_sync(0xffffffffu, accl, o);
      }
      __syncthreads();
      if (slot == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sh[w][sub * VEC + i] = acc[i];
        if (sub == 0) sh[w][32 * 4] = accl;
      }
      __syncthreads();
      // the chunk's (or the whole segment's) sum: warps combined in warp order
      float* out_row = chunked ? g_seg_part + (int64_t)it * D : (gsum != nullptr ? gsum + s * (int64_t)D : nullptr);
      float* out_lin = chunked ? g_seg_part_lin + it : (gsum_lin != nullptr ? gsum_lin + s : nullptr);
      if (w == 0 && slot == 0) {
        if (active && out_row != nullptr && (chunked ? demb != nullptr : true)) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float t = 0.f;
            for (int ww = 0; ww < 8; ++ww) t += sh[ww][sub * VEC + i];
            out_row[sub * VEC + i] = t;
          }
        }
        if (sub == 0 && out_lin != nullptr) {
          float t = 0.f;
          for (int ww = 0; ww < 8; ++ww) t += sh[ww][32 * 4];
          *out_lin = t;
        }
      }
      if (chunked) {
        // last chunk of the segment to finish adds the partial rows in chunk order (fixed order whoever is last)
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) sh_last = (atomicAdd(&g_seg_done[item.w], 1) == item.z - 1) ? 1 : 0;
        __syncthreads();
        if (sh_last) {
          __threadfence();
          if (gsum != nullptr && demb != nullptr) {
            for (int c0 = threadIdx.x; c0 < D; c0 += blockDim.x) {
              float t = 0.f;
              for (int c = 0; c < item.z; ++c) t += __ldcg(g_seg_part + (int64_t)(item.w + c) * D + c0);
              gsum[s * (int64_t)D + c0] = t;
            }
          }
          if (gsum_lin != nullptr && threadIdx.x == 0) {
            float t = 0.f;
            for (int c = 0; c < item.z; ++c) t += __ldcg(g_seg_part_lin + item.w + c);
            gsum_lin[s] = t;
          }
        }
      }
    }
  }
}

extern "C" int xdfm_embed_bwd_reduce(const float* demb, const float* dlin, const int32_t* sorted_pos, const int32_t* seg_offsets,
                                     const int32_t* num_segments, int64_t n_keys, int m, int D, float* gsum, float* gsum_lin,
                                     void* stream) {
  XDFM_CHECK_ARG(D >= 1 && D <= 128, "embed_bwd_reduce: D=%d unsupported (1..128)", D);
  if (n_keys == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int vec = (D % 4 == 0 && (uintptr_t)demb % 16 == 0) ? 4 : ((D % 2 == 0 && (uintptr_t)demb % 8 == 0) ? 2 : 1);
  int pieces = D / vec;
  int lpr = 1;
  while (lpr < pieces) lpr <<= 1;
  if (lpr > 32) {  // D > 32*vec cannot happen for vec=4 (D<=128); fall back to narrower rows never needed
    xdfm_set_error("embed_bwd_reduce: D=%d with vec=%d needs more than one warp per row", D, vec);
    return XDFM_ERR_UNSUPPORTED;
  }
  int blocks_short = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n_keys, 8));
  // chunks of all long segments: at most n_keys / SEG_CHUNK full ones + one partial per long segment (<= n_keys / SEG_LONG of those)
  const int64_t item_bound = n_keys / SEG_CHUNK + n_keys / SEG_LONG + 1;
  const bool chunked = item_bound <= SEG_ITEMS_MAX && item_bound * D <= SEG_PART_FLOATS;
  int blocks_long = chunked ? (int)min((int64_t)xdfm_num_sms() * 8, item_bound) : xdfm_num_sms() * 2;
  if (chunked) {
    void* cnt = nullptr;
    XDFM_CUDA(cudaGetSymbolAddress(&cnt, g_seg_item_count));
    XDFM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
  }
  const int mode_short = chunked ? 0 : -1, mode_long = chunked ? 3 : 2;
#define LAUNCH_SEG(V)                                                                                                    \
  seg_reduce_kernel<V><<<max(blocks_short, 1), 256, 0, st>>>(demb, dlin, sorted_pos, seg_offsets, num_segments, m, D, lpr, gsum, \
                                                             gsum_lin, mode_short);                                      \
  seg_reduce_kernel<V><<<blocks_long, 256, 0, st>>>(demb, dlin, sorted_pos, seg_offsets, num_segments, m, D, lpr, gsum,  \
                                                    gsum_lin, mode_long);
  if (vec == 4) { LAUNCH_SEG(4) } else if (vec == 2) { LAUNCH_SEG(2) } else { LAUNCH_SEG(1) }
#undef LAUNCH_SEG
  ++g_xdfm_launches;  // two kernels above, one check below
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// scatter the per-segment sums into dense per-table gradient tensors (generic/compat mode: user optimizers
// expect dense .grad like the reference's nn.Embedding(sparse=False)); grads must be pre-zeroed by the caller.
struct TablePtrsRW {
  float* p[XDFM_MAX_FIELDS];
  int64_t off[XDFM_MAX_FIELDS + 1];
};

__device__ __forceinline__ int find_table(const int64_t* off, int T, int64_t key) {
  int t = 0;
  while (t + 1 < T && key >= off[t + 1]) ++t;
  return t;
}

__global__ void scatter_dense_kernel(TablePtrsRW tp, int T, int width, const uint32_t* __restrict__ uniq_keys,
                                     const float* __restrict__ gsum, const int32_t* __restrict__ num_segments, float scale) {
  int nseg = *num_segments;
  int64_t total = (int64_t)nseg * width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i / width;
    int c = (int)(i - s * width);
    int64_t key = uniq_keys[s];
    int t = find_table(tp.off, T, key);
    tp.p[t][(key - tp.off[t]) * width + c] += scale * gsum[i];
  }
}

extern "C" int xdfm_embed_bwd_scatter_dense(float* const* grad_tables, const int64_t* table_row_offset, int T, int width,
                                            const uint32_t* uniq_keys, const float* gsum, const int32_t* num_segments,
                                            int64_t max_segments, void* stream) {
  XDFM_CHECK_ARG(T >= 1 && T <= XDFM_MAX_FIELDS, "scatter_dense: T=%d", T);
  if (max_segments == 0) return XDFM_OK;
  TablePtrsRW tp;
  for (int t = 0; t < T; ++t) { tp.p[t] = grad_tables[t]; tp.off[t] = table_row_offset[t]; }
  tp.off[T] = table_row_offset[T];
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(max_segments * width, 256));
  scatter_dense_kernel<<<max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(tp, T, width, uniq_keys, gsum, num_segments, 1.0f);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
