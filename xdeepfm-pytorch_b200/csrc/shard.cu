// Single-node multi-GPU embedding path: tables row-sharded across the GPUs of one NVSwitch domain, looked up and updated
// through PEER MEMORY (NVLink loads inside the kernels) instead of a pack -> all-to-all -> unpack pipeline.
//
// The reference has no equivalent (it replicates whole tables under nn.DataParallel: deepctr/models/basemodel.py:206-209,
// deepctr/inputs.py:167-180); this is the B200 replacement for that replication.
//
// Layout.  Global row r of table t lives on rank  r % G  at local row  r / G  of that rank's shard of table t.  Every rank keeps
// its shards of all tables of a table set in ONE contiguous buffer [rows_g, width] allocated with cudaMalloc and exported with
// cudaIpc*: rank g, feature f -> rows  feat_base[g*m + f] + id / G  of buffer g.  All ranks map all buffers, so
//   forward   out[b, f, :] = shard[id % G][feat_base[..] + id / G, :]         one kernel, remote rows come over NVLink
//   backward  each rank sorts ITS batch's keys  owner*S + local_row  (owner-major), reduces duplicate rows deterministically
//             and leaves (unique keys, row sums, per-owner ranges) in an exported exchange buffer; after a stream-ordered
//             barrier every owner PULLS its ranges from all peers (coalesced NVLink reads), merges them with a second
//             stable sort (source-rank order fixed -> bit-reproducible) and applies the fused optimizer to its shard.
// No variable-size collective, no host synchronisation: counts stay on the device.
#include "common.cuh"
#include <string.h>
#include <cub/cub.cuh>
#include "../../include/xdfm.h"

#define XDFM_MAX_RANKS 16

// ------------------------------------------------------------------------------------------------
// peer-mappable memory
// ------------------------------------------------------------------------------------------------
extern "C" int xdfm_ipc_alloc(int64_t bytes, void** dptr) {
  XDFM_CHECK_ARG(bytes > 0 && dptr != nullptr, "ipc_alloc: bytes=%lld", (long long)bytes);
  XDFM_CUDA(cudaMalloc(dptr, (size_t)bytes));
  XDFM_CUDA(cudaMemset(*dptr, 0, (size_t)bytes));
  return XDFM_OK;
}
extern "C" int xdfm_ipc_free(void* dptr) {
  if (dptr != nullptr) XDFM_CUDA(cudaFree(dptr));
  return XDFM_OK;
}
extern "C" int xdfm_ipc_export(void* dptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  XDFM_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), dptr));
  return XDFM_OK;
}
extern "C" int xdfm_ipc_open(const void* handle64, void** dptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  XDFM_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return XDFM_OK;
}
extern "C" int xdfm_ipc_close(void* dptr) {
  if (dptr != nullptr) XDFM_CUDA(cudaIpcCloseMemHandle(dptr));
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// forward: sharded gather (+ first-order term)
// ------------------------------------------------------------------------------------------------
struct VocabArr {
  int32_t v[XDFM_MAX_FIELDS];
};

__global__ void __launch_bounds__(256) embed_gather_sharded_kernel(const float* const* __restrict__ shards,
                                                                   const int64_t* __restrict__ feat_base, VocabArr vocab,
                                                                   const int32_t* __restrict__ ids, int64_t n_rows, int m, int D, int G,
                                                                   float* __restrict__ out) {
  const int vpr = D >> 2;   // float4 per row
  const int64_t total = n_rows * vpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    float4 val[U];
    int64_t dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      dst[u] = -1;
      if (i < total) {
        const int64_t row = i / vpr;
        const int v = (int)(i - row * vpr);
        const int f = (int)(row % m);
        int id = __ldg(ids + row);
        id = max(0, min(id, vocab.v[f] - 1));
        const int owner = id % G;
        const int64_t lrow = feat_base[owner * m + f] + id / G;
        const float4* src = reinterpret_cast<const float4*>(shards[owner] + lrow * D) + v;
        val[u] = *src;     // plain (coherent) load: the row may belong to a peer and was written by its optimizer kernel
        dst[u] = i;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] >= 0) reinterpret_cast<float4*>(out)[dst[u]] = val[u];
  }
}

__global__ void __launch_bounds__(256) linear_term_sharded_kernel(const float* const* __restrict__ lin_shards,
                                                                  const int64_t* __restrict__ feat_base, VocabArr vocab,
                                                                  const int32_t* __restrict__ ids, int64_t B, int m, int G,
                                                                  const float* __restrict__ dense, int nd,
                                                                  const float* __restrict__ dense_w, float* __restrict__ out_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float acc = 0.f;
    for (int f = lane; f < m; f += 32) {
      int id = __ldg(ids + b * m + f);
      id = max(0, min(id, vocab.v[f] - 1));
      const int owner = id % G;
      acc += *(lin_shards[owner] + feat_base[owner * m + f] + id / G);
    }
    float accd = 0.f;
    if (dense_w != nullptr)
      for (int j = lane; j < nd; j += 32) accd += __ldg(dense + b * nd + j) * __ldg(dense_w + j);
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) out_lin[b] = acc + accd;
  }
}

extern "C" int xdfm_embed_gather_sharded(const float* const* emb_shards_dev, const float* const* lin_shards_dev,
                                         const int64_t* feat_base_dev, const int32_t* vocab, const int32_t* ids, int64_t B, int m, int D,
                                         int G, float* out_emb, const float* dense, int nd, const float* dense_w, float* out_lin,
                                         void* stream) {
  XDFM_CHECK_ARG(m >= 1 && m <= XDFM_MAX_FIELDS, "embed_gather_sharded: m=%d out of range (max %d)", m, XDFM_MAX_FIELDS);
  XDFM_CHECK_ARG(G >= 1 && G <= XDFM_MAX_RANKS, "embed_gather_sharded: G=%d out of range (max %d)", G, XDFM_MAX_RANKS);
  XDFM_CHECK_ARG(feat_base_dev != nullptr, "embed_gather_sharded: feat_base is null");
  if (B == 0) return XDFM_OK;
  VocabArr va;
  for (int f = 0; f < m; ++f) {
    XDFM_CHECK_ARG(vocab[f] > 0, "embed_gather_sharded: vocab[%d]=%d", f, vocab[f]);
    va.v[f] = vocab[f];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (out_emb != nullptr) {
    XDFM_CHECK_ARG(emb_shards_dev != nullptr, "embed_gather_sharded: emb_shards is null");
    XDFM_CHECK_ARG(D >= 4 && D % 4 == 0 && (uintptr_t)out_emb % 16 == 0, "embed_gather_sharded: D=%d must be a multiple of 4", D);
    const int64_t total = B * (int64_t)m * (D / 4);
    int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(total, 256 * 4));
    embed_gather_sharded_kernel<<<max(blocks, 1), 256, 0, st>>>(emb_shards_dev, feat_base_dev, va, ids, B * (int64_t)m, m, D, G, out_emb);
    XDFM_LAUNCH_CHECK();
  }
  if (out_lin != nullptr) {
    XDFM_CHECK_ARG(lin_shards_dev != nullptr, "embed_gather_sharded: lin_shards is null");
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(B, 8));
    linear_term_sharded_kernel<<<max(blocks, 1), 256, 0, st>>>(lin_shards_dev, feat_base_dev, va, ids, B, m, G, dense, nd,
                                                               nd > 0 ? dense_w : nullptr, out_lin);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// sort -> run-length segments of an arbitrary uint32 key array (optionally with a trailing sentinel run that is dropped)
// ------------------------------------------------------------------------------------------------
static size_t align256s(size_t x) { return (x + 255) & ~(size_t)255; }

struct KeyWs {
  size_t keys_in, pos_in, counts, cub_temp, cub_bytes, total;
};

static KeyWs key_ws_layout(int64_t n) {
  KeyWs L;
  size_t o = 0;
  L.keys_in = o; o += align256s(n * sizeof(uint32_t));
  L.pos_in = o; o += align256s(n * sizeof(int32_t));
  L.counts = o; o += align256s((n + 1) * sizeof(int32_t));
  size_t sort_b = 0, rle_b = 0, scan_b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, 32);
  cub::DeviceRunLengthEncode::Encode(nullptr, rle_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                     (int32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  L.cub_bytes = align256s(max(sort_b, max(rle_b, scan_b)));
  L.cub_temp = o; o += L.cub_bytes;
  L.total = o;
  return L;
}

extern "C" int64_t xdfm_shard_workspace_bytes(int64_t n_keys) {
  if (n_keys <= 0) return 256;
  return (int64_t)key_ws_layout(n_keys).total;
}

__global__ void iota_kernel(int32_t* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

// seg_offsets[num] = n_valid; if the last run is the sentinel it is dropped from the count.
// owner_ranges[g] = first segment whose key >= g*S (g = 0..G), so rank g's segments are [ranges[g], ranges[g+1]).
__global__ void finalize_segments_kernel(int32_t* seg_offsets, int32_t* num_segments, const uint32_t* __restrict__ uniq_keys, int32_t n,
                                         uint32_t sentinel, int has_sentinel, int32_t* owner_ranges, int G, uint32_t S) {
  __shared__ int nseg_s;
  if (threadIdx.x == 0) {
    int nseg = *num_segments;
    if (has_sentinel && nseg > 0 && uniq_keys[nseg - 1] == sentinel) {
      --nseg;                       // seg_offsets[nseg] already is the start of the sentinel run = number of valid keys
      *num_segments = nseg;
    } else {
      seg_offsets[nseg] = n;
    }
    nseg_s = nseg;
  }
  __syncthreads();
  if (owner_ranges != nullptr && (int)threadIdx.x <= G) {
    const int nseg = nseg_s;
    const uint64_t target = (uint64_t)threadIdx.x * S;
    int lo = 0, hi = nseg;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((uint64_t)uniq_keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    owner_ranges[threadIdx.x] = lo;
  }
}

// keys_in/pos_in live at the head of the workspace and are already filled on `st`.
static int segments_from_ws(const KeyWs& L, char* ws, int64_t n, int end_bit, uint32_t* uniq_keys, int32_t* seg_offsets, int32_t* sorted_pos,
                            int32_t* num_segments, uint32_t sentinel, int has_sentinel, int32_t* owner_ranges, int G, uint32_t S,
                            cudaStream_t st) {
  uint32_t* keys_in = (uint32_t*)(ws + L.keys_in);
  int32_t* pos_in = (int32_t*)(ws + L.pos_in);
  int32_t* counts = (int32_t*)(ws + L.counts);
  void* cub_temp = ws + L.cub_temp;
  size_t tb = L.cub_bytes;
  XDFM_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, tb, (const uint32_t*)keys_in, uniq_keys, (const int32_t*)pos_in, sorted_pos, (int)n,
                                            0, end_bit, st));
  XDFM_CUDA(cudaMemcpyAsync(keys_in, uniq_keys, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  tb = L.cub_bytes;
  XDFM_CUDA(cub::DeviceRunLengthEncode::Encode(cub_temp, tb, (const uint32_t*)keys_in, uniq_keys, counts, num_segments, (int)n, st));
  tb = L.cub_bytes;
  XDFM_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp, tb, (const int32_t*)counts, seg_offsets, (int)n, st));
  finalize_segments_kernel<<<1, 32, 0, st>>>(seg_offsets, num_segments, uniq_keys, (int32_t)n, sentinel, has_sentinel, owner_ranges, G, S);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

__global__ void make_shard_keys_kernel(const int32_t* __restrict__ ids, int64_t n, int m, int G, uint32_t S,
                                       const int64_t* __restrict__ feat_base, VocabArr vocab, uint32_t* __restrict__ keys,
                                       int32_t* __restrict__ pos) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % m);
    const int id = max(0, min(ids[i], vocab.v[f] - 1));
    const int owner = id % G;
    keys[i] = (uint32_t)owner * S + (uint32_t)(feat_base[owner * m + f] + id / G);
    pos[i] = (int32_t)i;
  }
}

extern "C" int xdfm_shard_segments(const int32_t* ids, int64_t B, int m, int G, uint32_t key_stride, const int64_t* feat_base_dev,
                                   const int32_t* vocab, void* workspace, int64_t workspace_bytes, uint32_t* uniq_keys,
                                   int32_t* seg_offsets, int32_t* sorted_pos, int32_t* num_segments, int32_t* owner_ranges, void* stream) {
  XDFM_CHECK_ARG(m >= 1 && m <= XDFM_MAX_FIELDS, "shard_segments: m=%d", m);
  XDFM_CHECK_ARG(G >= 1 && G <= XDFM_MAX_RANKS, "shard_segments: G=%d", G);
  XDFM_CHECK_ARG(key_stride > 0 && (uint64_t)key_stride * (uint64_t)G <= ((uint64_t)1 << 32),
                 "shard_segments: G*key_stride = %llu exceeds the 32-bit key space", (unsigned long long)key_stride * G);
  const int64_t n = B * (int64_t)m;
  XDFM_CHECK_ARG(n < ((int64_t)1 << 31), "shard_segments: B*m too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    XDFM_CUDA(cudaMemsetAsync(num_segments, 0, sizeof(int32_t), st));
    if (owner_ranges != nullptr) XDFM_CUDA(cudaMemsetAsync(owner_ranges, 0, sizeof(int32_t) * (G + 1), st));
    return XDFM_OK;
  }
  KeyWs L = key_ws_layout(n);
  XDFM_CHECK_ARG(workspace_bytes >= (int64_t)L.total, "shard_segments: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                 (long long)L.total);
  VocabArr va;
  for (int f = 0; f < m; ++f) va.v[f] = vocab[f];
  char* ws = (char*)workspace;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  make_shard_keys_kernel<<<blocks, 256, 0, st>>>(ids, n, m, G, key_stride, feat_base_dev, va, (uint32_t*)(ws + L.keys_in),
                                                 (int32_t*)(ws + L.pos_in));
  XDFM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((uint64_t)1 << end_bit) < (uint64_t)key_stride * G) ++end_bit;
  return segments_from_ws(L, ws, n, end_bit, uniq_keys, seg_offsets, sorted_pos, num_segments, 0u, 0, owner_ranges, G, key_stride, st);
}

// ------------------------------------------------------------------------------------------------
// owner side: pull this rank's key ranges + row sums from every peer's exchange buffer (NVLink reads)
// ------------------------------------------------------------------------------------------------
struct PeerBufs {
  const uint32_t* keys[XDFM_MAX_RANKS];
  const float* gsum[XDFM_MAX_RANKS];
  const float* gsum_lin[XDFM_MAX_RANKS];
  const int32_t* ranges[XDFM_MAX_RANKS];
};

__global__ void __launch_bounds__(256) shard_pull_kernel(PeerBufs pb, int G, int rank, uint32_t S, int D, int64_t n_cap, uint32_t sentinel,
                                                         uint32_t* __restrict__ keys_out, int32_t* __restrict__ pos_out,
                                                         float* __restrict__ rows_out, float* __restrict__ rows_lin_out) {
  __shared__ int beg[XDFM_MAX_RANKS], pre[XDFM_MAX_RANKS + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int g = 0; g < G; ++g) {
      const int b = pb.ranges[g][rank], e = pb.ranges[g][rank + 1];
      beg[g] = b;
      pre[g] = acc;
      acc += e - b;
    }
    pre[G] = acc;
  }
  __syncthreads();
  const int64_t total = min((int64_t)pre[G], n_cap);
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  // keys (+ identity positions); the tail is filled with the sentinel so that a fixed-size sort can follow
  for (int64_t i = tid; i < n_cap; i += nthr) {
    uint32_t k = sentinel;
    if (i < total) {
      int g = 0;
      while (g + 1 < G && i >= pre[g + 1]) ++g;
      const int64_t src = beg[g] + (i - pre[g]);
      k = pb.keys[g][src] - (uint32_t)rank * S;
      if (rows_lin_out != nullptr) rows_lin_out[i] = pb.gsum_lin[g][src];
    }
    keys_out[i] = k;
    pos_out[i] = (int32_t)i;
  }
  if (rows_out != nullptr) {
    const int vpr = D >> 2;
    for (int64_t i = tid; i < total * vpr; i += nthr) {
      const int64_t r = i / vpr;
      const int v = (int)(i - r * vpr);
      int g = 0;
      while (g + 1 < G && r >= pre[g + 1]) ++g;
      const int64_t src = beg[g] + (r - pre[g]);
      reinterpret_cast<float4*>(rows_out)[i] = reinterpret_cast<const float4*>(pb.gsum[g] + src * D)[v];
    }
  }
}

extern "C" int xdfm_shard_pull_segments(const void* const* peer_keys, const void* const* peer_gsum, const void* const* peer_gsum_lin,
                                        const void* const* peer_ranges, int G, int rank, uint32_t key_stride, int D, int64_t n_cap,
                                        void* workspace, int64_t workspace_bytes, float* rows, float* rows_lin, uint32_t* uniq_keys,
                                        int32_t* seg_offsets, int32_t* sorted_pos, int32_t* num_segments, void* stream) {
  XDFM_CHECK_ARG(G >= 1 && G <= XDFM_MAX_RANKS && rank >= 0 && rank < G, "shard_pull: G=%d rank=%d", G, rank);
  XDFM_CHECK_ARG(n_cap > 0 && n_cap < ((int64_t)1 << 31), "shard_pull: n_cap=%lld", (long long)n_cap);
  XDFM_CHECK_ARG(rows == nullptr || (D % 4 == 0 && (uintptr_t)rows % 16 == 0), "shard_pull: D=%d must be a multiple of 4", D);
  KeyWs L = key_ws_layout(n_cap);
  XDFM_CHECK_ARG(workspace_bytes >= (int64_t)L.total, "shard_pull: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                 (long long)L.total);
  PeerBufs pb;
  for (int g = 0; g < G; ++g) {
    pb.keys[g] = (const uint32_t*)peer_keys[g];
    pb.gsum[g] = peer_gsum ? (const float*)peer_gsum[g] : nullptr;
    pb.gsum_lin[g] = peer_gsum_lin ? (const float*)peer_gsum_lin[g] : nullptr;
    pb.ranges[g] = (const int32_t*)peer_ranges[g];
    XDFM_CHECK_ARG(pb.keys[g] != nullptr && pb.ranges[g] != nullptr, "shard_pull: peer %d buffers are null", g);
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const uint32_t sentinel = key_stride;   // one past the largest local key
  int blocks = xdfm_num_sms() * 8;
  shard_pull_kernel<<<blocks, 256, 0, st>>>(pb, G, rank, key_stride, D, n_cap, sentinel, (uint32_t*)(ws + L.keys_in),
                                            (int32_t*)(ws + L.pos_in), peer_gsum ? rows : nullptr, peer_gsum_lin ? rows_lin : nullptr);
  XDFM_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && ((uint64_t)1 << end_bit) <= (uint64_t)sentinel) ++end_bit;
  return segments_from_ws(L, ws, n_cap, end_bit, uniq_keys, seg_offsets, sorted_pos, num_segments, sentinel, 1, nullptr, 0, 0, st);
}
