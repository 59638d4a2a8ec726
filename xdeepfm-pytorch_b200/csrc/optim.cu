// Fused optimizer + L2-regulariser kernels.
//
// Replaces (reference, file:line):
//   get_regularization_loss  deepctr/models/basemodel.py:412-428  (sum(l2 * p^2) over every registered tensor,
//                                                                 including WHOLE embedding tables, every step)
//   optim.step()             deepctr/models/basemodel.py:262, 447-461 (torch.optim.SGD/Adam/Adagrad/RMSprop, torch
//                                                                 defaults, dense over every table row)
// Semantics are the reference's: every row of every table is updated every step with
// g = scatter_add(row grads) + 2*l2*w, so the pass is a pure HBM stream: 4 B read + 4 B write per state
// element (Adam: w, m, v -> 24 B/element).  Rows touched by the batch are updated by the sparse kernel
// (which also marks them in a bitmap); the dense kernel streams all tables and skips marked rows.
#include <cstring>

#include "common.cuh"
#include "../../include/xdfm.h"

// device-side per-step scalars written by opt_tick_kernel (keeps the step CUDA-graph capturable)
// d[0] = step (int32 bits), d[1] = adam step_size, d[2] = adam 1/sqrt(bias_correction2), d[3] = adagrad clr
__global__ void opt_tick_kernel(float* d, xdfm_opt_cfg cfg) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int step = __float_as_int(d[0]) + 1;
  d[0] = __int_as_float(step);
  double bc1 = 1.0 - pow((double)cfg.beta1, (double)step);
  double bc2 = 1.0 - pow((double)cfg.beta2, (double)step);
  d[1] = (float)((double)cfg.lr / bc1);
  d[2] = (float)(1.0 / sqrt(bc2));
  d[3] = (float)((double)cfg.lr / (1.0 + (double)(step - 1) * (double)cfg.lr_decay));
}

extern "C" int xdfm_opt_tick(float* opt_dev, const xdfm_opt_cfg* cfg, void* stream) {
  opt_tick_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(opt_dev, *cfg);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

struct OptScalars {
  int kind;
  float lr, one_minus_b1, b2, one_minus_b2, eps, alpha, one_minus_alpha, step_size, bc2_sqrt, clr;   // bc2_sqrt holds 1/sqrt(bc2)
};

__device__ __forceinline__ OptScalars load_scalars(const xdfm_opt_cfg& c, const float* __restrict__ d) {
  OptScalars s;
  s.kind = c.kind;
  s.lr = c.lr;
  s.one_minus_b1 = (float)(1.0 - (double)c.beta1);
  s.b2 = c.beta2;
  s.one_minus_b2 = (float)(1.0 - (double)c.beta2);
  s.eps = c.eps;
  s.alpha = c.alpha;
  s.one_minus_alpha = (float)(1.0 - (double)c.alpha);
  s.step_size = d[1];
  s.bc2_sqrt = d[2];
  s.clr = d[3];
  return s;
}

// one element; mirrors torch's single-tensor optimizer arithmetic (torch/optim/{sgd,adam,adagrad,rmsprop}.py).
// Every operation is an explicit intrinsic: the compiler may not contract or re-associate, so the streaming pass, the touched-row
// kernel and the lazy replay produce bit-identical results wherever the same update is computed.  Square roots and quotients use
// the SFU approximations (sqrt.approx / rcp-based division, <= 2 ulp): the lazy replay of postponed rows is pure ALU work and two
// IEEE divisions per element and step would triple its cost; the parity tests against torch.optim hold at 1e-4 of the update.
// (.ftz forms: ONE MUFU instruction for the square root, MUFU.RCP + FMUL for the quotient; without .ftz each carries a
// subnormal-range check and rescaling, ~4 extra instructions, i.e. half of the replay's issue slots.  Subnormal inputs are flushed:
// that only concerns second moments below 1e-38, where sqrt(v) + eps == eps either way.)
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_div(float a, float b) {
  float r;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void opt_apply(const OptScalars& h, float& w, float g, float& s1, float& s2) {
  switch (h.kind) {
    case XDFM_OPT_SGD:
      w = __fsub_rn(w, __fmul_rn(h.lr, g));
      break;
    case XDFM_OPT_ADAM: {
      s1 = __fmaf_rn(h.one_minus_b1, __fsub_rn(g, s1), s1);                      // exp_avg.lerp_(grad, 1-beta1)
      s2 = __fmaf_rn(h.one_minus_b2, __fmul_rn(g, g), __fmul_rn(s2, h.b2));      // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
      const float denom = __fmaf_rn(fast_sqrt(s2), h.bc2_sqrt, h.eps);           // sqrt(v) / sqrt(bias_correction2) + eps
      w = __fsub_rn(w, __fmul_rn(h.step_size, fast_div(s1, denom)));           // param.addcdiv_(exp_avg, denom, value=-step_size)
      break;
    }
    case XDFM_OPT_ADAGRAD: {
      s1 = __fmaf_rn(g, g, s1);                                                  // state_sum.addcmul_(grad, grad, value=1)
      const float stdv = __fadd_rn(fast_sqrt(s1), h.eps);
      w = __fsub_rn(w, __fmul_rn(h.clr, fast_div(g, stdv)));
      break;
    }
    case XDFM_OPT_RMSPROP: {
      s1 = __fmaf_rn(h.one_minus_alpha, __fmul_rn(g, g), __fmul_rn(s1, h.alpha));   // square_avg.mul_(alpha).addcmul_(g, g, 1-alpha)
      const float avg = __fadd_rn(fast_sqrt(s1), h.eps);
      w = __fsub_rn(w, __fmul_rn(h.lr, fast_div(g, avg)));
      break;
    }
  }
}
// gradient of the L2 term alone (rows the batch did not touch) / added to a scattered gradient
__device__ __forceinline__ float l2_grad(float l2, float w) { return __fmul_rn(__fmul_rn(2.f, l2), w); }
__device__ __forceinline__ float l2_plus_grad(float gsum, float grad_scale, float l2, float w) {
  return __fadd_rn(__fmul_rn(gsum, grad_scale), l2_grad(l2, w));
}

__device__ __forceinline__ void block_accumulate_double(float local, double* out) {
  __shared__ float red[32];
  float v = warp_sum(local);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0 && out != nullptr) atomicAdd(out, (double)t);
  }
}

// ------------------------------------------------------------------------------------------------
// flat dense-parameter step: w, g, s1, s2 are flat fp32 buffers over ALL dense parameters
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flat_opt_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, int64_t n, float* __restrict__ w,
                                                       const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2,
                                                       const float* __restrict__ l2vec, float grad_scale, double* reg_out) {
  OptScalars h = load_scalars(cfg, d);
  float reg = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float wi = w[i];
    float l2 = l2vec ? l2vec[i] : 0.f;
    reg += l2 * (wi * wi);
    float gi = l2_plus_grad(g[i], grad_scale, l2, wi);
    float a = s1 ? s1[i] : 0.f, b = s2 ? s2[i] : 0.f;
    opt_apply(h, wi, gi, a, b);
    w[i] = wi;
    if (s1) s1[i] = a;
    if (s2) s2[i] = b;
  }
  block_accumulate_double(reg, reg_out);
}

extern "C" int xdfm_flat_opt(const xdfm_opt_cfg* cfg, const float* opt_dev, int64_t n, float* w, const float* g, float* s1,
                             float* s2, const float* l2vec, float grad_scale, double* reg_out, void* stream) {
  XDFM_CHECK_ARG(cfg->kind >= 0 && cfg->kind <= 3, "flat_opt: unknown optimizer kind %d", cfg->kind);
  if (n == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  flat_opt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*cfg, opt_dev, n, w, g, s1, s2, l2vec, grad_scale, reg_out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// embedding tables: sparse (touched rows) + dense (all other rows)
// ------------------------------------------------------------------------------------------------
struct TableSet {
  float* w[XDFM_MAX_FIELDS];
  float* s1[XDFM_MAX_FIELDS];
  float* s2[XDFM_MAX_FIELDS];
  int64_t row_off[XDFM_MAX_FIELDS + 1];   // global row offsets (key space of the scatter-add)
  int64_t vec_off[XDFM_MAX_FIELDS + 1];   // offsets in units of 4 elements, each table padded to a multiple of 4
};

__device__ __forceinline__ int find_tab(const int64_t* off, int T, int64_t key) {
  int lo = 0, hi = T - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (key >= off[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) rows_opt_sparse_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, TableSet ts, int T, int width,
                                                              const uint32_t* __restrict__ uniq_keys, const float* __restrict__ gsum,
                                                              const int32_t* __restrict__ num_segments, float grad_scale,
                                                              uint32_t* __restrict__ touched, double* reg_out) {
  OptScalars h = load_scalars(cfg, d);
  const int nseg = *num_segments;
  const int64_t total = (int64_t)nseg * width;
  float reg = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i / width;
    int c = (int)(i - s * width);
    int64_t key = uniq_keys[s];
    int t = find_tab(ts.row_off, T, key);
    int64_t e = (key - ts.row_off[t]) * width + c;
    float wi = ts.w[t][e];
    reg += cfg.l2 * (wi * wi);
    float gi = l2_plus_grad(gsum[i], grad_scale, cfg.l2, wi);
    float a = ts.s1[t] ? ts.s1[t][e] : 0.f, b = ts.s2[t] ? ts.s2[t][e] : 0.f;
    opt_apply(h, wi, gi, a, b);
    ts.w[t][e] = wi;
    if (ts.s1[t]) ts.s1[t][e] = a;
    if (ts.s2[t]) ts.s2[t][e] = b;
    if (c == 0 && touched != nullptr) atomicOr(touched + (key >> 5), 1u << (key & 31));
  }
  block_accumulate_double(reg, reg_out);
}

// dense pass over every row NOT marked in `touched`: g = 2*l2*w.  One thread = 4 consecutive elements of one table.
__global__ void __launch_bounds__(256) rows_opt_dense_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, TableSet ts, int T, int width,
                                                             const uint32_t* __restrict__ touched, double* reg_out) {
  OptScalars h = load_scalars(cfg, d);
  const int64_t nvec = ts.vec_off[T];
  float reg = 0.f;
  const bool has1 = ts.s1[0] != nullptr, has2 = ts.s2[0] != nullptr;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    int t = find_tab(ts.vec_off, T, v);
    int64_t e0 = (v - ts.vec_off[t]) * 4;
    int64_t nelem = (ts.row_off[t + 1] - ts.row_off[t]) * width;
    float wv[4], av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    bool full = e0 + 4 <= nelem;
    if (full) {
      *reinterpret_cast<float4*>(wv) = *reinterpret_cast<const float4*>(ts.w[t] + e0);
      if (has1) *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(ts.s1[t] + e0);
      if (has2) *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(ts.s2[t] + e0);
    } else {
      for (int i = 0; i < 4; ++i) {
        if (e0 + i < nelem) {
          wv[i] = ts.w[t][e0 + i];
          if (has1) av[i] = ts.s1[t][e0 + i];
          if (has2) bv[i] = ts.s2[t][e0 + i];
        }
      }
    }
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t e = e0 + i;
      if (e < nelem) {
        int64_t key = ts.row_off[t] + e / width;
        bool tch = touched != nullptr && ((touched[key >> 5] >> (key & 31)) & 1u);
        if (!tch) {
          float wi = wv[i];
          reg += cfg.l2 * (wi * wi);
          opt_apply(h, wv[i], l2_grad(cfg.l2, wi), av[i], bv[i]);
          any = true;
        }
      }
    }
    if (!any) continue;
    if (full) {
      *reinterpret_cast<float4*>(ts.w[t] + e0) = *reinterpret_cast<float4*>(wv);
      if (has1) *reinterpret_cast<float4*>(ts.s1[t] + e0) = *reinterpret_cast<float4*>(av);
      if (has2) *reinterpret_cast<float4*>(ts.s2[t] + e0) = *reinterpret_cast<float4*>(bv);
    } else {
      for (int i = 0; i < 4; ++i) {
        if (e0 + i < nelem) {
          ts.w[t][e0 + i] = wv[i];
          if (has1) ts.s1[t][e0 + i] = av[i];
          if (has2) ts.s2[t][e0 + i] = bv[i];
        }
      }
    }
  }
  block_accumulate_double(reg, reg_out);
}

// v2 of the dense pass: two independent 4-element vectors per thread and iteration (6 x 16-byte loads in flight before any
// arithmetic), streaming (evict-first) loads/stores -- every byte is touched exactly once per step --, one bitmap probe per
// vector when a row is a whole number of vectors, optimizer kind resolved at compile time.
__device__ __forceinline__ float4 ld_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }

template <int KIND>
__device__ __forceinline__ void opt_apply_k(const OptScalars& h, float& w, float g, float& s1, float& s2) {
  OptScalars hk = h;
  hk.kind = KIND;            // compile-time constant: the switch in opt_apply folds away
  opt_apply(hk, w, g, s1, s2);
}

template <int KIND>
__global__ void __launch_bounds__(256, 4) rows_opt_dense_v2_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, TableSet ts, int T, int width,
                                                                   const uint32_t* __restrict__ touched, double* reg_out) {
  constexpr bool HAS1 = KIND != XDFM_OPT_SGD, HAS2 = KIND == XDFM_OPT_ADAM;
  constexpr int U = 2;
  OptScalars h = load_scalars(cfg, d);
  const int64_t nvec = ts.vec_off[T];
  const int wv = width >> 2;                 // vectors per row (width % 4 == 0 on this path)
  float reg = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * U) {
    float4 wq[U], aq[U], bq[U];
    float* pw[U];
    float* pa[U];
    float* pb[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      live[u] = v < nvec;
      if (live[u]) {
        const int t = find_tab(ts.vec_off, T, v);
        const int64_t lv = v - ts.vec_off[t];
        const int64_t key = ts.row_off[t] + lv / wv;
        pw[u] = ts.w[t] + lv * 4;
        wq[u] = ld_stream(pw[u]);
        if (HAS1) { pa[u] = ts.s1[t] + lv * 4; aq[u] = ld_stream(pa[u]); }
        if (HAS2) { pb[u] = ts.s2[t] + lv * 4; bq[u] = ld_stream(pb[u]); }
        if (touched != nullptr && ((__ldg(touched + (key >> 5)) >> (key & 31)) & 1u)) live[u] = false;   // updated by the sparse kernel
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      float* wf = reinterpret_cast<float*>(&wq[u]);
      float* af = reinterpret_cast<float*>(&aq[u]);
      float* bf = reinterpret_cast<float*>(&bq[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wi = wf[i];
        reg += cfg.l2 * (wi * wi);
        float a = HAS1 ? af[i] : 0.f, b = HAS2 ? bf[i] : 0.f;
        opt_apply_k<KIND>(h, wf[i], l2_grad(cfg.l2, wi), a, b);
        if (HAS1) af[i] = a;
        if (HAS2) bf[i] = b;
      }
      st_stream(pw[u], wq[u]);
      if (HAS1) st_stream(pa[u], aq[u]);
      if (HAS2) st_stream(pb[u], bq[u]);
    }
  }
  block_accumulate_double(reg, reg_out);
}

int g_rows_opt_dense_version = 2;
extern "C" void xdfm_set_rows_opt_dense_version(int v) { g_rows_opt_dense_version = v; }

static int fill_table_set(TableSet& ts, float* const* w, float* const* s1, float* const* s2, const int64_t* row_off, int T, int width) {
  int64_t voff = 0;
  for (int t = 0; t < T; ++t) {
    ts.w[t] = w[t];
    ts.s1[t] = s1 ? s1[t] : nullptr;
    ts.s2[t] = s2 ? s2[t] : nullptr;
    ts.row_off[t] = row_off[t];
    ts.vec_off[t] = voff;
    voff += ((row_off[t + 1] - row_off[t]) * width + 3) / 4;
    if ((uintptr_t)w[t] % 16 != 0) return 1;
  }
  ts.row_off[T] = row_off[T];
  ts.vec_off[T] = voff;
  return 0;
}

extern "C" int xdfm_rows_opt(const xdfm_opt_cfg* cfg, const float* opt_dev, float* const* w, float* const* s1, float* const* s2,
                             const int64_t* table_row_offset, int T, int width, const uint32_t* uniq_keys, const float* gsum,
                             const int32_t* num_segments, int64_t max_segments, float grad_scale, uint32_t* touched_bitmap,
                             int dense_pass, double* reg_out, void* stream) {
  XDFM_CHECK_ARG(cfg->kind >= 0 && cfg->kind <= 3, "rows_opt: unknown optimizer kind %d", cfg->kind);
  XDFM_CHECK_ARG(T >= 1 && T <= XDFM_MAX_FIELDS, "rows_opt: T=%d", T);
  TableSet ts;
  XDFM_CHECK_ARG(fill_table_set(ts, w, s1, s2, table_row_offset, T, width) == 0, "rows_opt: tables must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total_rows = table_row_offset[T];
  if (dense_pass && touched_bitmap != nullptr)
    XDFM_CUDA(cudaMemsetAsync(touched_bitmap, 0, (size_t)((total_rows + 31) / 32) * 4, st));
  if (max_segments > 0 && uniq_keys != nullptr) {
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(max_segments * width, 256));
    rows_opt_sparse_kernel<<<max(blocks, 1), 256, 0, st>>>(*cfg, opt_dev, ts, T, width, uniq_keys, gsum, num_segments, grad_scale,
                                                           dense_pass ? touched_bitmap : nullptr, reg_out);
    XDFM_LAUNCH_CHECK();
  }
  if (dense_pass) {
    int64_t nvec = ts.vec_off[T];
    if (nvec > 0) {
      bool whole = width % 4 == 0;     // every table is a whole number of 4-element vectors and every vector lies inside one row
      if (g_rows_opt_dense_version == 2 && whole) {
        int blocks = (int)min((int64_t)xdfm_num_sms() * 4 * 4, ceil_div64(nvec, 256 * 2));
        blocks = max(blocks, 1);
        switch (cfg->kind) {
          case XDFM_OPT_SGD: rows_opt_dense_v2_kernel<XDFM_OPT_SGD><<<blocks, 256, 0, st>>>(*cfg, opt_dev, ts, T, width, touched_bitmap, reg_out); break;
          case XDFM_OPT_ADAM: rows_opt_dense_v2_kernel<XDFM_OPT_ADAM><<<blocks, 256, 0, st>>>(*cfg, opt_dev, ts, T, width, touched_bitmap, reg_out); break;
          case XDFM_OPT_ADAGRAD: rows_opt_dense_v2_kernel<XDFM_OPT_ADAGRAD><<<blocks, 256, 0, st>>>(*cfg, opt_dev, ts, T, width, touched_bitmap, reg_out); break;
          default: rows_opt_dense_v2_kernel<XDFM_OPT_RMSPROP><<<blocks, 256, 0, st>>>(*cfg, opt_dev, ts, T, width, touched_bitmap, reg_out); break;
        }
      } else {
        int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(nvec, 256));
        rows_opt_dense_kernel<<<max(blocks, 1), 256, 0, st>>>(*cfg, opt_dev, ts, T, width, touched_bitmap, reg_out);
      }
      XDFM_LAUNCH_CHECK();
    }
  }
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// Lazy ("deferred catch-up") form of the reference's dense table semantics.
//
// The reference moves EVERY table row EVERY step (g = 2*l2*w through Adam's moments: basemodel.py:126, 412-428, 447-461), which
// costs a 24 B/element stream over all tables per step (13.8 GB at Criteo cardinalities) although a batch touches ~0.6 % of the
// rows.  A row that is not touched evolves autonomously -- its update depends only on its own (w, m, v) and on the step's bias
// corrections -- so the update can be postponed without changing a single bit: `last[row]` records the step up to which the row
// is current, every step's scalars are kept in `hist`, and whoever needs the row (the batch that looks it up, a flush before
// predict / state_dict / the end of an epoch) replays the missing steps in registers with exactly the arithmetic of the dense
// pass.  HBM traffic drops to the touched rows; the replay is pure ALU work.
//   hist[(s - hist_base) * 4 + {0,1,2}] = (adam step_size, adam sqrt(bias_correction2), adagrad clr) of step s.
// ------------------------------------------------------------------------------------------------
__global__ void opt_tick_hist_kernel(float* d, xdfm_opt_cfg cfg, float* hist, long long hist_cap, long long hist_base) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int step = __float_as_int(d[0]) + 1;
  d[0] = __int_as_float(step);
  double bc1 = 1.0 - pow((double)cfg.beta1, (double)step);
  double bc2 = 1.0 - pow((double)cfg.beta2, (double)step);
  d[1] = (float)((double)cfg.lr / bc1);
  d[2] = (float)(1.0 / sqrt(bc2));
  d[3] = (float)((double)cfg.lr / (1.0 + (double)(step - 1) * (double)cfg.lr_decay));
  long long slot = (long long)step - hist_base;
  if (hist != nullptr && slot >= 0 && slot < hist_cap) {
    hist[slot * 4 + 0] = d[1];
    hist[slot * 4 + 1] = d[2];
    hist[slot * 4 + 2] = d[3];
    hist[slot * 4 + 3] = cfg.lr;
  }
}

extern "C" int xdfm_opt_tick_hist(float* opt_dev, const xdfm_opt_cfg* cfg, float* hist, int64_t hist_cap, int64_t hist_base, void* stream) {
  opt_tick_hist_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(opt_dev, *cfg, hist, (long long)hist_cap, (long long)hist_base);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// diagnostic switch (tests): 0 = scalar replay everywhere
__device__ int g_replay_packed = 1;
extern "C" int xdfm_set_replay_packed(int on) {
  XDFM_CUDA(cudaMemcpyToSymbol(g_replay_packed, &on, sizeof(int)));
  return XDFM_OK;
}

// Packed fp32 pairs (Blackwell fma / mul / add / sub .f32x2): each lane is the IEEE round-to-nearest result of the scalar intrinsic,
// so a replay written with them agrees bit for bit with opt_apply -- at half the issue slots.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// a - b as fma(b, -1, a): exact product, one rounding == sub.rn.  (ptxas contracts `mul.rn.f32x2` followed by `sub.rn.f32x2` into
// one FFMA2 -- the explicit .rn does not protect the packed forms the way it protects the scalar ones -- which changed 2 % of the
// replayed weights by an ulp or more against the scalar kernels.  A product cannot be folded into an FMA's operand.)
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
  const f32x2_t neg1 = 0xbf800000bf800000ull;                       // (-1.f, -1.f)
  return fma2(b, neg1, a);
}

// Adam replay of a 4-element piece.  The replay of postponed rows is pure ALU work (every element of every table, every step the
// row was not looked up): ~20 issue slots per element and step in the scalar form, next to two MUFU operations (sqrt.approx, the
// reciprocal of the quotient) that cannot be avoided.  Here the multiply / add work runs on packed pairs; same bits as opt_apply.
__device__ __forceinline__ void replay_adam4(const xdfm_opt_cfg& cfg, const OptScalars& h, const float4* __restrict__ hist, long long hist_base,
                                             int from, int to, float* w, float* a, float* b, float& reg) {
  const float c2 = __fmul_rn(2.f, cfg.l2);
  const f32x2_t c2l2 = pk2(c2, c2), l2p = pk2(cfg.l2, cfg.l2);
  const f32x2_t omb1 = pk2(h.one_minus_b1, h.one_minus_b1), omb2 = pk2(h.one_minus_b2, h.one_minus_b2), b2p = pk2(h.b2, h.b2);
  const f32x2_t epsp = pk2(h.eps, h.eps);
  f32x2_t W[2] = {pk2(w[0], w[1]), pk2(w[2], w[3])}, A[2] = {pk2(a[0], a[1]), pk2(a[2], a[3])}, Bv[2] = {pk2(b[0], b[1]), pk2(b[2], b[3])};
  f32x2_t regp = pk2(0.f, 0.f);
  for (int s = from + 1; s <= to; ++s) {
    const float4 hs = __ldg(hist + ((long long)s - hist_base));
    const f32x2_t ss = pk2(hs.x, hs.x), bc = pk2(hs.y, hs.y);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      regp = fma2(l2p, mul2(W[p], W[p]), regp);
      const f32x2_t g = mul2(c2l2, W[p]);                                     // l2_grad
      A[p] = fma2(omb1, sub2(g, A[p]), A[p]);                                 // exp_avg.lerp_(grad, 1-beta1)
      Bv[p] = fma2(omb2, mul2(g, g), mul2(Bv[p], b2p));                       // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
      float v0, v1, m0, m1, d0, d1;
      unpk2(Bv[p], v0, v1);
      const f32x2_t denom = fma2(pk2(fast_sqrt(v0), fast_sqrt(v1)), bc, epsp);   // sqrt(v) / sqrt(bias_correction2) + eps
      unpk2(denom, d0, d1);
      unpk2(A[p], m0, m1);
      W[p] = sub2(W[p], mul2(ss, pk2(fast_div(m0, d0), fast_div(m1, d1))));   // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
  }
  unpk2(W[0], w[0], w[1]); unpk2(W[1], w[2], w[3]);
  unpk2(A[0], a[0], a[1]); unpk2(A[1], a[2], a[3]);
  unpk2(Bv[0], b[0], b[1]); unpk2(Bv[1], b[2], b[3]);
  float r0, r1;
  unpk2(regp, r0, r1);
  reg += r0 + r1;
}

// replay steps (from, to] of an untouched row piece: g = 2*l2*w at every step, exactly as rows_opt_dense does
template <int VEC>
__device__ __forceinline__ void replay_steps(const xdfm_opt_cfg& cfg, OptScalars h, const float4* __restrict__ hist, long long hist_base,
                                             int from, int to, float* w, float* a, float* b, float& reg) {
  if (VEC == 4 && cfg.kind == XDFM_OPT_ADAM && g_replay_packed) {
    replay_adam4(cfg, h, hist, hist_base, from, to, w, a, b, reg);
    return;
  }
  for (int s = from + 1; s <= to; ++s) {
    const float4 hs = __ldg(hist + ((long long)s - hist_base));
    h.step_size = hs.x;
    h.bc2_sqrt = hs.y;
    h.clr = hs.z;
    h.lr = hs.w;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float wi = w[i];
      reg += cfg.l2 * (wi * wi);
      opt_apply(h, w[i], l2_grad(cfg.l2, wi), a[i], b[i]);
    }
  }
}

// rows named by uniq_keys: replay up to `done` completed steps (done = step counter in d[0]) and write back.
// One thread = one 4-element piece of a row (width % 4 == 0) or one element (generic).
template <int VEC>
__global__ void __launch_bounds__(256) rows_catchup_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, const float4* __restrict__ hist,
                                                           long long hist_base, TableSet ts, int T, int width, int32_t* __restrict__ last,
                                                           const uint32_t* __restrict__ uniq_keys, const int32_t* __restrict__ num_segments,
                                                           double* reg_out) {
  OptScalars h = load_scalars(cfg, d);
  const int done = __float_as_int(d[0]);
  const int nseg = *num_segments;
  const int ppr = width / VEC;                       // pieces per row
  const int64_t total = (int64_t)nseg * ppr;
  const bool has1 = ts.s1[0] != nullptr, has2 = ts.s2[0] != nullptr;
  float reg = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t sidx = i / ppr;
    const int piece = (int)(i - sidx * ppr);
    const int64_t key = uniq_keys[sidx];
    const int from = last[key];
    if (from >= done) continue;
    const int t = find_tab(ts.row_off, T, key);
    const int64_t e = (key - ts.row_off[t]) * width + (int64_t)piece * VEC;
    float w[VEC], a[VEC], b[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      w[k] = ts.w[t][e + k];
      a[k] = has1 ? ts.s1[t][e + k] : 0.f;
      b[k] = has2 ? ts.s2[t][e + k] : 0.f;
    }
    replay_steps<VEC>(cfg, h, hist, hist_base, from, done, w, a, b, reg);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      ts.w[t][e + k] = w[k];
      if (has1) ts.s1[t][e + k] = a[k];
      if (has2) ts.s2[t][e + k] = b[k];
    }
  }
  block_accumulate_double(reg, reg_out);
}

// second pass (separate launch: all pieces of a row must have read last[key] before it changes)
__global__ void rows_mark_kernel(int32_t* __restrict__ last, const uint32_t* __restrict__ uniq_keys, const int32_t* __restrict__ num_segments,
                                 const float* __restrict__ d, int plus) {
  const int nseg = *num_segments;
  const int v = __float_as_int(d[0]) + plus;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nseg; i += (int64_t)gridDim.x * blockDim.x) last[uniq_keys[i]] = v;
}

// every row: replay up to `done`; then last[:] = done (second launch)
template <int VEC>
__global__ void __launch_bounds__(256) rows_flush_kernel(xdfm_opt_cfg cfg, const float* __restrict__ d, const float4* __restrict__ hist,
                                                         long long hist_base, TableSet ts, int T, int width, const int32_t* __restrict__ last,
                                                         double* reg_out) {
  OptScalars h = load_scalars(cfg, d);
  const int done = __float_as_int(d[0]);
  const int64_t npieces = ts.vec_off[T];             // VEC == 4: 4-element vectors; VEC == 1: elements (vec_off built accordingly)
  const bool has1 = ts.s1[0] != nullptr, has2 = ts.s2[0] != nullptr;
  float reg = 0.f;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < npieces; v += (int64_t)gridDim.x * blockDim.x) {
    const int t = find_tab(ts.vec_off, T, v);
    const int64_t e = (v - ts.vec_off[t]) * VEC;
    const int64_t key = ts.row_off[t] + e / width;
    const int from = last[key];
    if (from >= done) continue;
    float w[VEC], a[VEC], b[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      w[k] = ts.w[t][e + k];
      a[k] = has1 ? ts.s1[t][e + k] : 0.f;
      b[k] = has2 ? ts.s2[t][e + k] : 0.f;
    }
    replay_steps<VEC>(cfg, h, hist, hist_base, from, done, w, a, b, reg);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      ts.w[t][e + k] = w[k];
      if (has1) ts.s1[t][e + k] = a[k];
      if (has2) ts.s2[t][e + k] = b[k];
    }
  }
  block_accumulate_double(reg, reg_out);
}

__global__ void fill_i32_from_step_kernel(int32_t* __restrict__ last, int64_t n, const float* __restrict__ d) {
  const int v = __float_as_int(d[0]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) last[i] = v;
}

static int fill_table_set_pieces(TableSet& ts, float* const* w, float* const* s1, float* const* s2, const int64_t* row_off, int T, int width,
                                 int vec) {
  int64_t voff = 0;
  for (int t = 0; t < T; ++t) {
    ts.w[t] = w[t];
    ts.s1[t] = s1 ? s1[t] : nullptr;
    ts.s2[t] = s2 ? s2[t] : nullptr;
    ts.row_off[t] = row_off[t];
    ts.vec_off[t] = voff;
    voff += (row_off[t + 1] - row_off[t]) * width / vec;
  }
  ts.row_off[T] = row_off[T];
  ts.vec_off[T] = voff;
  return 0;
}

extern "C" int xdfm_rows_catchup(const xdfm_opt_cfg* cfg, const float* opt_dev, const float* hist, int64_t hist_base, float* const* w,
                                 float* const* s1, float* const* s2, int32_t* last, const int64_t* table_row_offset, int T, int width,
                                 const uint32_t* uniq_keys, const int32_t* num_segments, int64_t max_segments, double* reg_out, void* stream) {
  XDFM_CHECK_ARG(cfg->kind >= 0 && cfg->kind <= 3, "rows_catchup: unknown optimizer kind %d", cfg->kind);
  XDFM_CHECK_ARG(T >= 1 && T <= XDFM_MAX_FIELDS && hist != nullptr && last != nullptr, "rows_catchup: bad arguments (T=%d)", T);
  if (max_segments == 0) return XDFM_OK;
  const int vec = width % 4 == 0 ? 4 : 1;
  TableSet ts;
  fill_table_set_pieces(ts, w, s1, s2, table_row_offset, T, width, vec);
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(max_segments * (width / vec), 256));
  blocks = max(blocks, 1);
  if (vec == 4)
    rows_catchup_kernel<4><<<blocks, 256, 0, st>>>(*cfg, opt_dev, (const float4*)hist, (long long)hist_base, ts, T, width, last, uniq_keys,
                                                   num_segments, reg_out);
  else
    rows_catchup_kernel<1><<<blocks, 256, 0, st>>>(*cfg, opt_dev, (const float4*)hist, (long long)hist_base, ts, T, width, last, uniq_keys,
                                                   num_segments, reg_out);
  XDFM_LAUNCH_CHECK();
  int mblocks = (int)min((int64_t)xdfm_num_sms() * 4, ceil_div64(max_segments, 256));
  rows_mark_kernel<<<max(mblocks, 1), 256, 0, st>>>(last, uniq_keys, num_segments, opt_dev, 0);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// after the optimizer step of the touched rows (xdfm_rows_opt with dense_pass = 0): last[row] = step
extern "C" int xdfm_rows_mark_current(int32_t* last, const uint32_t* uniq_keys, const int32_t* num_segments, int64_t max_segments,
                                      const float* opt_dev, void* stream) {
  if (max_segments == 0) return XDFM_OK;
  int mblocks = (int)min((int64_t)xdfm_num_sms() * 4, ceil_div64(max_segments, 256));
  rows_mark_kernel<<<max(mblocks, 1), 256, 0, (cudaStream_t)stream>>>(last, uniq_keys, num_segments, opt_dev, 0);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_rows_flush(const xdfm_opt_cfg* cfg, const float* opt_dev, const float* hist, int64_t hist_base, float* const* w,
                               float* const* s1, float* const* s2, int32_t* last, const int64_t* table_row_offset, int T, int width,
                               double* reg_out, void* stream) {
  XDFM_CHECK_ARG(cfg->kind >= 0 && cfg->kind <= 3, "rows_flush: unknown optimizer kind %d", cfg->kind);
  XDFM_CHECK_ARG(T >= 1 && T <= XDFM_MAX_FIELDS && hist != nullptr && last != nullptr, "rows_flush: bad arguments (T=%d)", T);
  const int vec = width % 4 == 0 ? 4 : 1;
  TableSet ts;
  fill_table_set_pieces(ts, w, s1, s2, table_row_offset, T, width, vec);
  const int64_t np = ts.vec_off[T];
  if (np == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(np, 256));
  if (vec == 4)
    rows_flush_kernel<4><<<max(blocks, 1), 256, 0, st>>>(*cfg, opt_dev, (const float4*)hist, (long long)hist_base, ts, T, width, last, reg_out);
  else
    rows_flush_kernel<1><<<max(blocks, 1), 256, 0, st>>>(*cfg, opt_dev, (const float4*)hist, (long long)hist_base, ts, T, width, last, reg_out);
  XDFM_LAUNCH_CHECK();
  const int64_t rows = table_row_offset[T];
  int fb = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(rows, 256));
  fill_i32_from_step_kernel<<<max(fb, 1), 256, 0, st>>>(last, rows, opt_dev);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// Lazy semantics for ROW-SHARDED tables (csrc/shard.cu): the reader replays.  A rank that looks up a row owned by a peer reads
// (w, m, v, last) through peer memory and, if the row is behind, replays the missing steps in registers (read-only: nothing is
// written back, the owner does its own catch-up when it applies the step's update).  All ranks run the same steps with the same
// hyper-parameters, so every rank's local `hist` is identical.
//   ptrs: DEVICE int64 [8 * G]: for g in 0..G-1: emb, lin, s1, s1_lin, s2, s2_lin, last, last_lin (peer-mapped addresses of rank g's shard)
// ------------------------------------------------------------------------------------------------
struct VocabArr2 {
  int32_t v[XDFM_MAX_FIELDS];
};

__global__ void __launch_bounds__(256) gather_sharded_lazy_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ feat_base,
                                                                  VocabArr2 vocab, const int32_t* __restrict__ ids, int64_t n_rows, int m,
                                                                  int D, int G, xdfm_opt_cfg cfg, const float* __restrict__ d,
                                                                  const float4* __restrict__ hist, long long hist_base,
                                                                  float* __restrict__ out) {
  OptScalars h = load_scalars(cfg, d);
  const int done = __float_as_int(d[0]);
  const int vpr = D >> 2;
  const int64_t total = n_rows * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vpr;
    const int v = (int)(i - row * vpr);
    const int f = (int)(row % m);
    int id = __ldg(ids + row);
    id = max(0, min(id, vocab.v[f] - 1));
    const int owner = id % G;
    const int64_t lrow = feat_base[owner * m + f] + id / G;
    const long long* pp = ptrs + owner * 8;
    const int64_t e = lrow * D + v * 4;
    // every operand of the replay is requested at once: a remote row costs ONE NVLink round trip, not two (the moments used to be
    // fetched only after `last` had come back; nearly every row a batch looks up is behind, so nothing is saved by waiting)
    float4 w4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[0]) + e);
    const int from = *(reinterpret_cast<const int32_t*>(pp[6]) + lrow);
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pp[2]) a4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[2]) + e);
    if (pp[4]) b4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[4]) + e);
    if (from < done) {
      float w[4] = {w4.x, w4.y, w4.z, w4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
      float reg = 0.f;
      replay_steps<4>(cfg, h, hist, hist_base, from, done, w, a, b, reg);
      w4 = make_float4(w[0], w[1], w[2], w[3]);
    }
    reinterpret_cast<float4*>(out)[i] = w4;
  }
}

__global__ void __launch_bounds__(256) linear_term_sharded_lazy_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ feat_base,
                                                                       VocabArr2 vocab, const int32_t* __restrict__ ids, int64_t B, int m, int G,
                                                                       xdfm_opt_cfg cfg, const float* __restrict__ d,
                                                                       const float4* __restrict__ hist, long long hist_base,
                                                                       const float* __restrict__ dense, int nd, const float* __restrict__ dense_w,
                                                                       float* __restrict__ out_lin) {
  OptScalars h = load_scalars(cfg, d);
  const int done = __float_as_int(d[0]);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float acc = 0.f;
    for (int f = lane; f < m; f += 32) {
      int id = __ldg(ids + b * m + f);
      id = max(0, min(id, vocab.v[f] - 1));
      const int owner = id % G;
      const int64_t lrow = feat_base[owner * m + f] + id / G;
      const long long* pp = ptrs + owner * 8;
      float w[1] = {*(reinterpret_cast<const float*>(pp[1]) + lrow)};
      const int from = *(reinterpret_cast<const int32_t*>(pp[7]) + lrow);
      float a[1] = {pp[3] ? *(reinterpret_cast<const float*>(pp[3]) + lrow) : 0.f};      // requested with w and `last`: one round trip
      float bb[1] = {pp[5] ? *(reinterpret_cast<const float*>(pp[5]) + lrow) : 0.f};
      if (from < done) {
        float reg = 0.f;
        replay_steps<1>(cfg, h, hist, hist_base, from, done, w, a, bb, reg);
      }
      acc += w[0];
    }
    float accd = 0.f;
    if (dense_w != nullptr)
      for (int j = lane; j < nd; j += 32) accd += __ldg(dense + b * nd + j) * __ldg(dense_w + j);
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) out_lin[b] = acc + accd;
  }
}

// ------------------------------------------------------------------------------------------------
// Row-sharded lookup through the batch's UNIQUE rows.  A Criteo-shaped batch looks most rows up many times (cfg2: 213 k lookups,
// 54 k distinct rows), and a remote row costs four small NVLink reads (w, both moments, `last`): the direct kernels above are bound
// by the request rate of the links, not by bytes.  Here every distinct row crosses NVLink once -- keys, segments and sorted positions
// are the ones the backward needs anyway (xdfm_shard_segments, now run before the lookup) -- is replayed if stale, and lands in a
// compact local buffer; the per-sample tensors are expanded from it with local loads.
//   uniq_keys[s] = owner * key_stride + local row;  inv[q] = segment of lookup q = b * m + f
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fetch_unique_emb_kernel(const long long* __restrict__ ptrs, const uint32_t* __restrict__ uniq_keys,
                                                               const int32_t* __restrict__ num_segments, uint32_t S, int D, xdfm_opt_cfg cfg,
                                                               const float* __restrict__ d, const float4* __restrict__ hist,
                                                               long long hist_base, float* __restrict__ u_emb) {
  const bool lazy = hist != nullptr;
  OptScalars h = lazy ? load_scalars(cfg, d) : OptScalars();
  const int done = lazy ? __float_as_int(d[0]) : 0;
  const int vpr = D >> 2;
  const int64_t total = (int64_t)(*num_segments) * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / vpr;
    const int v = (int)(i - s * vpr);
    const uint32_t key = __ldg(uniq_keys + s);
    const int owner = (int)(key / S);
    const int64_t lrow = key - (uint32_t)owner * S;
    const long long* pp = ptrs + owner * 8;
    const int64_t e = lrow * D + v * 4;
    float4 w4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[0]) + e);
    if (lazy) {
      // every operand of the replay is requested at once: one NVLink round trip per remote row
      const int from = *(reinterpret_cast<const int32_t*>(pp[6]) + lrow);
      float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pp[2]) a4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[2]) + e);
      if (pp[4]) b4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp[4]) + e);
      if (from < done) {
        float w[4] = {w4.x, w4.y, w4.z, w4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
        float reg = 0.f;
        replay_steps<4>(cfg, h, hist, hist_base, from, done, w, a, b, reg);
        w4 = make_float4(w[0], w[1], w[2], w[3]);
      }
    }
    reinterpret_cast<float4*>(u_emb)[i] = w4;
  }
}

__global__ void __launch_bounds__(256) fetch_unique_lin_kernel(const long long* __restrict__ ptrs, const uint32_t* __restrict__ uniq_keys,
                                                               const int32_t* __restrict__ num_segments, uint32_t S, xdfm_opt_cfg cfg,
                                                               const float* __restrict__ d, const float4* __restrict__ hist,
                                                               long long hist_base, float* __restrict__ u_lin) {
  const bool lazy = hist != nullptr;
  OptScalars h = lazy ? load_scalars(cfg, d) : OptScalars();
  const int done = lazy ? __float_as_int(d[0]) : 0;
  const int nseg = *num_segments;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nseg; s += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = __ldg(uniq_keys + s);
    const int owner = (int)(key / S);
    const int64_t lrow = key - (uint32_t)owner * S;
    const long long* pp = ptrs + owner * 8;
    float w[1] = {*(reinterpret_cast<const float*>(pp[1]) + lrow)};
    if (lazy) {
      const int from = *(reinterpret_cast<const int32_t*>(pp[7]) + lrow);
      float a[1] = {pp[3] ? *(reinterpret_cast<const float*>(pp[3]) + lrow) : 0.f};
      float bb[1] = {pp[5] ? *(reinterpret_cast<const float*>(pp[5]) + lrow) : 0.f};
      if (from < done) {
        float reg = 0.f;
        replay_steps<1>(cfg, h, hist, hist_base, from, done, w, a, bb, reg);
      }
    }
    u_lin[s] = w[0];
  }
}

// inv[sorted_pos[p]] = the segment that holds sorted position p (binary search in seg_offsets[0 .. nseg])
__global__ void __launch_bounds__(256) segment_of_lookup_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ sorted_pos,
                                                                const int32_t* __restrict__ num_segments, int64_t n, int32_t* __restrict__ inv) {
  const int nseg = *num_segments;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = nseg;                          // last s with seg_offsets[s] <= p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(seg_offsets + mid) <= p) lo = mid; else hi = mid;
    }
    inv[sorted_pos[p]] = lo;
  }
}

__global__ void __launch_bounds__(256) expand_unique_emb_kernel(const float4* __restrict__ u_emb, const int32_t* __restrict__ inv, int64_t n,
                                                                int vpr, float4* __restrict__ out) {
  const int64_t total = n * vpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = i / vpr;
    const int v = (int)(i - q * vpr);
    out[i] = u_emb[(int64_t)__ldg(inv + q) * vpr + v];
  }
}

// same lane / field assignment and summation order as linear_term_sharded(_lazy)_kernel: bit-identical first-order term
__global__ void __launch_bounds__(256) expand_unique_lin_kernel(const float* __restrict__ u_lin, const int32_t* __restrict__ inv, int64_t B, int m,
                                                                const float* __restrict__ dense, int nd, const float* __restrict__ dense_w,
                                                                float* __restrict__ out_lin) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float acc = 0.f;
    for (int f = lane; f < m; f += 32) acc += u_lin[__ldg(inv + b * m + f)];
    float accd = 0.f;
    if (dense_w != nullptr)
      for (int j = lane; j < nd; j += 32) accd += __ldg(dense + b * nd + j) * __ldg(dense_w + j);
    acc = warp_sum(acc);
    accd = warp_sum(accd);
    if (lane == 0) out_lin[b] = acc + accd;
  }
}

// Distinct rows of the batch -> u_emb [nseg, D], u_lin [nseg] (stale rows replayed when hist != NULL; cfg_* / opt_dev may be NULL
// otherwise), and inv [n].  n = B * m = number of lookups (capacity of every per-lookup / per-segment array).
extern "C" int xdfm_embed_fetch_unique_sharded(const void* ptrs_dev, int G, uint32_t key_stride, int D, const uint32_t* uniq_keys,
                                               const int32_t* seg_offsets, const int32_t* sorted_pos, const int32_t* num_segments, int64_t n,
                                               const xdfm_opt_cfg* cfg_emb, const xdfm_opt_cfg* cfg_lin, const float* opt_dev, const float* hist,
                                               int64_t hist_base, float* u_emb, float* u_lin, int32_t* inv, void* stream) {
  XDFM_CHECK_ARG(G >= 1 && G <= 16 && key_stride > 0, "embed_fetch_unique_sharded: G=%d key_stride=%u", G, key_stride);
  XDFM_CHECK_ARG(ptrs_dev != nullptr && uniq_keys != nullptr && seg_offsets != nullptr && sorted_pos != nullptr && num_segments != nullptr,
                 "embed_fetch_unique_sharded: null argument");
  XDFM_CHECK_ARG(hist == nullptr || (opt_dev != nullptr && cfg_emb != nullptr && cfg_lin != nullptr),
                 "embed_fetch_unique_sharded: lazy tables need cfg_emb, cfg_lin and opt_dev");
  XDFM_CHECK_ARG(u_emb == nullptr || (D >= 4 && D % 4 == 0), "embed_fetch_unique_sharded: D=%d must be a multiple of 4", D);
  if (n == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  xdfm_opt_cfg none;
  memset(&none, 0, sizeof(none));
  if (u_emb != nullptr) {
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n * (D / 4), 256));
    fetch_unique_emb_kernel<<<max(blocks, 1), 256, 0, st>>>((const long long*)ptrs_dev, uniq_keys, num_segments, key_stride, D,
                                                            hist ? *cfg_emb : none, opt_dev, (const float4*)hist, (long long)hist_base, u_emb);
    XDFM_LAUNCH_CHECK();
  }
  if (u_lin != nullptr) {
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
    fetch_unique_lin_kernel<<<max(blocks, 1), 256, 0, st>>>((const long long*)ptrs_dev, uniq_keys, num_segments, key_stride,
                                                            hist ? *cfg_lin : none, opt_dev, (const float4*)hist, (long long)hist_base, u_lin);
    XDFM_LAUNCH_CHECK();
  }
  if (inv != nullptr) {
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
    segment_of_lookup_kernel<<<max(blocks, 1), 256, 0, st>>>(seg_offsets, sorted_pos, num_segments, n, inv);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

// first-order rows of every lookup, un-summed: out [n] = u_lin[inv] (models with multi-value features pool them per field first)
__global__ void __launch_bounds__(256) expand_unique_rows1_kernel(const float* __restrict__ u_lin, const int32_t* __restrict__ inv, int64_t n,
                                                                  float* __restrict__ out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) out[q] = u_lin[__ldg(inv + q)];
}

extern "C" int xdfm_embed_expand_unique_lin_rows(const float* u_lin, const int32_t* inv, int64_t n, float* out, void* stream) {
  XDFM_CHECK_ARG(n == 0 || (u_lin != nullptr && inv != nullptr && out != nullptr), "embed_expand_unique_lin_rows: null argument");
  if (n == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(n, 256));
  expand_unique_rows1_kernel<<<max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(u_lin, inv, n, out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// out_emb [B, m, D] = u_emb[inv] and / or out_lin [B] = sum_f u_lin[inv[b, f]] + dense[b, :] . dense_w (either output may be NULL)
extern "C" int xdfm_embed_expand_unique(const float* u_emb, const float* u_lin, const int32_t* inv, int64_t B, int m, int D, float* out_emb,
                                        const float* dense, int nd, const float* dense_w, float* out_lin, void* stream) {
  XDFM_CHECK_ARG(m >= 1 && m <= XDFM_MAX_FIELDS && inv != nullptr, "embed_expand_unique: m=%d", m);
  if (B == 0) return XDFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_emb != nullptr) {
    XDFM_CHECK_ARG(u_emb != nullptr && D >= 4 && D % 4 == 0 && (uintptr_t)out_emb % 16 == 0 && (uintptr_t)u_emb % 16 == 0,
                   "embed_expand_unique: D=%d must be a multiple of 4 and the buffers 16-byte aligned", D);
    const int64_t total = B * (int64_t)m * (D / 4);
    int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(total, 256));
    expand_unique_emb_kernel<<<max(blocks, 1), 256, 0, st>>>((const float4*)u_emb, inv, B * (int64_t)m, D / 4, (float4*)out_emb);
    XDFM_LAUNCH_CHECK();
  }
  if (out_lin != nullptr) {
    XDFM_CHECK_ARG(u_lin != nullptr, "embed_expand_unique: u_lin is null");
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(B, 8));
    expand_unique_lin_kernel<<<max(blocks, 1), 256, 0, st>>>(u_lin, inv, B, m, dense, nd, nd > 0 ? dense_w : nullptr, out_lin);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}

extern "C" int xdfm_embed_gather_sharded_lazy(const void* ptrs_dev, const int64_t* feat_base_dev, const int32_t* vocab, const int32_t* ids,
                                              int64_t B, int m, int D, int G, const xdfm_opt_cfg* cfg_emb, const xdfm_opt_cfg* cfg_lin,
                                              const float* opt_dev, const float* hist, int64_t hist_base, float* out_emb, const float* dense,
                                              int nd, const float* dense_w, float* out_lin, void* stream) {
  XDFM_CHECK_ARG(m >= 1 && m <= XDFM_MAX_FIELDS && G >= 1 && G <= 16, "embed_gather_sharded_lazy: m=%d G=%d", m, G);
  XDFM_CHECK_ARG(ptrs_dev != nullptr && feat_base_dev != nullptr && hist != nullptr && opt_dev != nullptr,
                 "embed_gather_sharded_lazy: null argument");
  if (B == 0) return XDFM_OK;
  VocabArr2 va;
  for (int f = 0; f < m; ++f) va.v[f] = vocab[f];
  cudaStream_t st = (cudaStream_t)stream;
  if (out_emb != nullptr) {
    XDFM_CHECK_ARG(D >= 4 && D % 4 == 0, "embed_gather_sharded_lazy: D=%d must be a multiple of 4", D);
    const int64_t total = B * (int64_t)m * (D / 4);
    int blocks = (int)min((int64_t)xdfm_num_sms() * 16, ceil_div64(total, 256));
    gather_sharded_lazy_kernel<<<max(blocks, 1), 256, 0, st>>>((const long long*)ptrs_dev, (const long long*)feat_base_dev, va, ids,
                                                               B * (int64_t)m, m, D, G, *cfg_emb, opt_dev, (const float4*)hist,
                                                               (long long)hist_base, out_emb);
    XDFM_LAUNCH_CHECK();
  }
  if (out_lin != nullptr) {
    int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(B, 8));
    linear_term_sharded_lazy_kernel<<<max(blocks, 1), 256, 0, st>>>((const long long*)ptrs_dev, (const long long*)feat_base_dev, va, ids, B, m,
                                                                    G, *cfg_lin, opt_dev, (const float4*)hist, (long long)hist_base, dense, nd,
                                                                    nd > 0 ? dense_w : nullptr, out_lin);
    XDFM_LAUNCH_CHECK();
  }
  return XDFM_OK;
}
