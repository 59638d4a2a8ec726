// Field self-attention block over the CIN feature maps (xDeepFMAttention / V2).
//
// Replaces (reference, file:line):
//   MultiHeadSelfAttention.forward: view to heads, softmax(Q K^T / sqrt(hd)) V      deepctr/layers/cin_attention.py:73-95
//   residual add + nn.LayerNorm(E)                                                   deepctr/layers/cin_attention.py:305-311, 455-460
//   AttentionPooling: softmax over L of the scores, weighted sum of the sequence     deepctr/layers/cin_attention.py:138-142
// (the bias-free E x E projections W_q/W_k/W_v/W_o and the Linear-Tanh-Linear score MLP run through xdfm_gemm_f32).
//
// The "sequence" is the L = featuremap_num CIN maps of ONE sample, the embedding dim is E = D (8..64), head_dim = E / heads is
// 2..16: QK^T has K = head_dim, far too small for tensor cores -- the block is exp/FMA-issue bound.  The
// reference materialises scores and probabilities [B, h, L, L] in HBM (1 MB/sample at L=256); here one CTA owns one sample,
// K and V (forward) / Q and dO (backward) live in shared memory, every thread owns one query (or key) row and streams over the
// other axis with an online softmax -- nothing of size L x L ever leaves the SM.  The backward recomputes the probabilities
// from the saved log-sum-exp (flash-attention style).
#include "common.cuh"
#include <math_constants.h>
#include "../../include/xdfm.h"

#define ATT_THREADS 256

__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int n) {
  if ((n & 3) == 0 && (((uintptr_t)src) & 15) == 0) {
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  }
}

// 2^x on the SFU without exp2f's denormal-range wrapper (two predicated multiplies and a compare per call): the arguments here are
// score differences <= 0 and results below 2^-126 contribute nothing to a softmax sum
__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One head's slice of a shared-memory row -> registers.  EXACT (head_dim == HDM): 128- / 64-bit loads, no predicates -- with
// predicated scalar loads the inner loops issued 8 LDS per key and the kernels were bound by the one-LDS-per-clock shared-memory
// pipe (round 1: 19.4 ms of the 29 ms BASELINE config-3 step).  Otherwise elements >= hd read as zero, so the arithmetic loops
// can always run over all HDM lanes.
template <int HDM, bool EXACT>
__device__ __forceinline__ void load_head(float (&dst)[HDM], const float* __restrict__ src, int hd) {
  if constexpr (EXACT && HDM % 4 == 0) {
#pragma unroll
    for (int i = 0; i < HDM / 4; ++i) {
      const float4 t = reinterpret_cast<const float4*>(src)[i];
      dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
    }
  } else if constexpr (EXACT && HDM == 2) {
    const float2 t = *reinterpret_cast<const float2*>(src);
    dst[0] = t.x; dst[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < HDM; ++i) dst[i] = i < hd ? src[i] : 0.f;
  }
}

// Dropout on the attention probabilities (reference: nn.Dropout(attn_probs), cin_attention.py:54, 86).  The probabilities never
// leave the SM, so the keep mask is a counter-based hash of (seed, sample, head, query, key) that the forward and both phases of the
// backward recompute: no mask tensor, no RNG state.  32-bit avalanche mixers (murmur3 / splitmix finalizers) over the 64-bit
// element index and the 64-bit seed; keep iff hash >= p * 2^32.  (Bit-wise parity with torch's Philox stream is not a goal:
// SURVEY.md section 7, hard part 8.)
struct DropCfg {
  const unsigned long long* seed;     // device [1]
  unsigned int threshold;             // p * 2^32 (p < 1)
  float inv_keep;                     // 1 / (1 - p)
};

__device__ __forceinline__ unsigned int mix32(unsigned int x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long row_index, int j, unsigned int threshold) {
  // row_index = ((b * nh + h) * L + l): one 64-bit value per (sample, head, query); j = key
  const unsigned int a = mix32((unsigned int)row_index ^ (unsigned int)seed);
  const unsigned int c = mix32((unsigned int)(row_index >> 32) + (unsigned int)(seed >> 32) * 0x9e3779b9u + a);
  return mix32(c ^ ((unsigned int)j * 0x9e3779b1u + 0x7f4a7c15u)) >= threshold;
}

// HDM = compile-time bound of head_dim; EXACT = (head_dim == HDM); DROP = dropout on the probabilities
template <int HDM, bool EXACT, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS) mhsa_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, int L, int E, int nh, int hd,
                                                               float scale_log2e, float* __restrict__ o, float* __restrict__ lse,
                                                               DropCfg drop) {
  extern __shared__ __align__(16) float sm_att[];
  const unsigned long long seed = DROP ? *drop.seed : 0ull;
  float* sK = sm_att;
  float* sV = sm_att + (size_t)L * E;
  const int64_t b = blockIdx.x;
  const size_t base = (size_t)b * L * E;
  stage_rows(sK, k + base, L * E);
  stage_rows(sV, v + base, L * E);
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    for (int h = 0; h < nh; ++h) {
      float qv[HDM], acc[HDM];
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        qv[i] = i < hd ? q[base + (size_t)l * E + h * hd + i] * scale_log2e : 0.f;
        acc[i] = 0.f;
      }
      float mx = -CUDART_INF_F, s = 0.f;
      const float* kr = sK + h * hd;
      const float* vr = sV + h * hd;
      const unsigned long long rix = ((unsigned long long)b * nh + h) * L + l;
      for (int j = 0; j < L; ++j, kr += E, vr += E) {
        float kk[HDM], vv[HDM];
        load_head<HDM, EXACT>(kk, kr, hd);
        load_head<HDM, EXACT>(vv, vr, hd);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < HDM; ++i) d = fmaf(qv[i], kk[i], d);
        if (d > mx) {
          const float c = fast_exp2(mx - d);
          s *= c;
#pragma unroll
          for (int i = 0; i < HDM; ++i) acc[i] *= c;
          mx = d;
        }
        const float p = fast_exp2(d - mx);
        s += p;                                   // the softmax denominator sees every key; dropped keys only leave the numerator
        const float pk = (!DROP || drop_keep(seed, rix, j, drop.threshold)) ? p : 0.f;
#pragma unroll
        for (int i = 0; i < HDM; ++i) acc[i] = fmaf(pk, vv[i], acc[i]);
      }
      const float inv = (DROP ? drop.inv_keep : 1.f) / s;
#pragma unroll
      for (int i = 0; i < HDM; ++i)
        if (i < hd) o[base + (size_t)l * E + h * hd + i] = acc[i] * inv;
      lse[((size_t)b * nh + h) * L + l] = mx + log2f(s);      // log2 units of the scaled scores
    }
  }
}

template <int HDM, bool EXACT, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS) mhsa_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                               const float* __restrict__ v, const float* __restrict__ o,
                                                               const float* __restrict__ lse, const float* __restrict__ dout, int L, int E,
                                                               int nh, int hd, float scale, float* __restrict__ dq, float* __restrict__ dk,
                                                               float* __restrict__ dv, DropCfg drop) {
  extern __shared__ __align__(16) float sm_att[];
  const unsigned long long seed = DROP ? *drop.seed : 0ull;
  // with P~ = keep / (1 - p) * P:  o = P~ V,  dP = keep / (1 - p) * (dO . v),  D = sum_j P dP = dO . o (unchanged),
  // dS = P * (dP - D),  dV_j = sum_l P~_lj dO_l
  float* sA = sm_att;                          // phase A: K      phase B: Q
  float* sB = sm_att + (size_t)L * E;          // phase A: V      phase B: dO
  float* sLse = sB + (size_t)L * E;            // [nh][L]
  float* sD = sLse + (size_t)L * nh;           // [nh][L]   D = rowsum(dO * O)
  const int64_t b = blockIdx.x;
  const size_t base = (size_t)b * L * E;
  const float scale_log2e = scale * 1.4426950408889634f;
  stage_rows(sA, k + base, L * E);
  stage_rows(sB, v + base, L * E);
  for (int i = threadIdx.x; i < L * nh; i += blockDim.x) sLse[i] = lse[(size_t)b * nh * L + i];
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    for (int h = 0; h < nh; ++h) {
      float d = 0.f;
      for (int i = 0; i < hd; ++i) d = fmaf(dout[base + (size_t)l * E + h * hd + i], o[base + (size_t)l * E + h * hd + i], d);
      sD[h * L + l] = d;
    }
  }
  __syncthreads();
  // ---- phase A: thread = query row l -> dQ
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    for (int h = 0; h < nh; ++h) {
      float qv[HDM], dov[HDM], acc[HDM];
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        qv[i] = i < hd ? q[base + (size_t)l * E + h * hd + i] * scale_log2e : 0.f;
        dov[i] = i < hd ? dout[base + (size_t)l * E + h * hd + i] : 0.f;
        acc[i] = 0.f;
      }
      const float ls = sLse[h * L + l], Dl = sD[h * L + l];
      const float* kr = sA + h * hd;
      const float* vr = sB + h * hd;
      const unsigned long long rix = ((unsigned long long)b * nh + h) * L + l;
      for (int j = 0; j < L; ++j, kr += E, vr += E) {
        float kk[HDM], vv[HDM];
        load_head<HDM, EXACT>(kk, kr, hd);
        load_head<HDM, EXACT>(vv, vr, hd);
        float d = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < HDM; ++i) { d = fmaf(qv[i], kk[i], d); dp = fmaf(dov[i], vv[i], dp); }
        if (DROP) dp = drop_keep(seed, rix, j, drop.threshold) ? dp * drop.inv_keep : 0.f;
        const float ds = fast_exp2(d - ls) * (dp - Dl);
#pragma unroll
        for (int i = 0; i < HDM; ++i) acc[i] = fmaf(ds, kk[i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < HDM; ++i)
        if (i < hd) dq[base + (size_t)l * E + h * hd + i] = acc[i] * scale;
    }
  }
  __syncthreads();
  // ---- phase B: thread = key row j -> dK, dV
  stage_rows(sA, q + base, L * E);
  stage_rows(sB, dout + base, L * E);
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    for (int h = 0; h < nh; ++h) {
      float kv[HDM], vv[HDM], dkv[HDM], dvv[HDM];
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        kv[i] = i < hd ? k[base + (size_t)j * E + h * hd + i] * scale_log2e : 0.f;
        vv[i] = i < hd ? v[base + (size_t)j * E + h * hd + i] : 0.f;
        dkv[i] = 0.f;
        dvv[i] = 0.f;
      }
      const float* qr = sA + h * hd;
      const float* dor = sB + h * hd;
      const float* lsr = sLse + h * L;
      const float* Dr = sD + h * L;
      for (int l = 0; l < L; ++l, qr += E, dor += E) {
        float qq[HDM], dd[HDM];
        load_head<HDM, EXACT>(qq, qr, hd);
        load_head<HDM, EXACT>(dd, dor, hd);
        float d = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < HDM; ++i) { d = fmaf(qq[i], kv[i], d); dp = fmaf(dd[i], vv[i], dp); }
        const float p = fast_exp2(d - lsr[l]);
        float pk = p;
        if (DROP) {
          const bool keep = drop_keep(seed, ((unsigned long long)b * nh + h) * L + l, j, drop.threshold);
          dp = keep ? dp * drop.inv_keep : 0.f;
          pk = keep ? p * drop.inv_keep : 0.f;
        }
        const float ds = p * (dp - Dr[l]);
#pragma unroll
        for (int i = 0; i < HDM; ++i) { dvv[i] = fmaf(pk, dd[i], dvv[i]); dkv[i] = fmaf(ds, qq[i], dkv[i]); }
      }
#pragma unroll
      for (int i = 0; i < HDM; ++i)
        if (i < hd) {
          dk[base + (size_t)j * E + h * hd + i] = dkv[i] * scale;
          dv[base + (size_t)j * E + h * hd + i] = dvv[i];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row-blocked variants (head_dim == HDM): a thread owns RB = 4 query rows (forward, backward phase A) or 4 key rows (phase B) of ONE
// head and streams over the other axis.  Every shared-memory row slice it loads (K / V, or Q / dO) is used for four rows:
// shared memory delivers 128 bytes per clock to the register file, broadcast or not, and with one row per thread the two
// 128-bit loads per step (8 clocks per warp) -- not the ~15 arithmetic instructions -- set the pace (round 1: forward 3.3 ms at
// BASELINE config 3).  Work items = heads x ceil(L / 4) row groups, rows of a group L/4 apart (lanes walk consecutive rows).
// ------------------------------------------------------------------------------------------------
#define ATT_RB 4

// Packed fp32 arithmetic (sm_100 FFMA2 / FMUL2: two IEEE fp32 operations per instruction).  These kernels are bound by instruction
// issue, and 8 of the ~13-21 instructions per (query, key) pair are multiply-adds over the head's HDM (even) lanes.
template <int HDM>
__device__ __forceinline__ float dot_p(const float (&a)[HDM], const float (&b)[HDM]) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < HDM; i += 2) acc = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(b[i], b[i + 1]), acc);
  return acc.x + acc.y;
}
template <int HDM>
__device__ __forceinline__ void axpy_p(float s, const float (&x)[HDM], float (&y)[HDM]) {      // y += s * x
  const float2 s2 = make_float2(s, s);
#pragma unroll
  for (int i = 0; i < HDM; i += 2) {
    const float2 r = __ffma2_rn(s2, make_float2(x[i], x[i + 1]), make_float2(y[i], y[i + 1]));
    y[i] = r.x;
    y[i + 1] = r.y;
  }
}
template <int HDM>
__device__ __forceinline__ void scale_p(float s, float (&y)[HDM]) {                              // y *= s
  const float2 s2 = make_float2(s, s);
#pragma unroll
  for (int i = 0; i < HDM; i += 2) {
    const float2 r = __fmul2_rn(s2, make_float2(y[i], y[i + 1]));
    y[i] = r.x;
    y[i + 1] = r.y;
  }
}

template <int HDM>
__global__ void __launch_bounds__(ATT_THREADS) mhsa_fwd_rb_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                  const float* __restrict__ v, int L, int E, int nh,
                                                                  float scale_log2e, float* __restrict__ o, float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm_att[];
  float* sK = sm_att;
  float* sV = sm_att + (size_t)L * E;
  const int64_t b = blockIdx.x;
  const size_t base = (size_t)b * L * E;
  stage_rows(sK, k + base, L * E);
  stage_rows(sV, v + base, L * E);
  __syncthreads();
  const int G = (L + ATT_RB - 1) / ATT_RB;
  for (int item = threadIdx.x; item < nh * G; item += blockDim.x) {
    const int h = item / G, g = item - h * G;
    float qv[ATT_RB][HDM], acc[ATT_RB][HDM], mx[ATT_RB], s[ATT_RB];
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int l = g + r * G;
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        qv[r][i] = l < L ? q[base + (size_t)l * E + h * HDM + i] * scale_log2e : 0.f;
        acc[r][i] = 0.f;
      }
      mx[r] = -CUDART_INF_F;
      s[r] = 0.f;
    }
    const float* kr = sK + h * HDM;
    const float* vr = sV + h * HDM;
    for (int j = 0; j < L; ++j, kr += E, vr += E) {
      float kk[HDM], vv[HDM];
      load_head<HDM, true>(kk, kr, HDM);
      load_head<HDM, true>(vv, vr, HDM);
#pragma unroll
      for (int r = 0; r < ATT_RB; ++r) {
        const float d = dot_p<HDM>(qv[r], kk);
        if (d > mx[r]) {
          const float c = fast_exp2(mx[r] - d);
          s[r] *= c;
          scale_p<HDM>(c, acc[r]);
          mx[r] = d;
        }
        const float p = fast_exp2(d - mx[r]);
        s[r] += p;
        axpy_p<HDM>(p, vv, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int l = g + r * G;
      if (l < L) {
        const float inv = 1.f / s[r];
#pragma unroll
        for (int i = 0; i < HDM; ++i) o[base + (size_t)l * E + h * HDM + i] = acc[r][i] * inv;
        lse[((size_t)b * nh + h) * L + l] = mx[r] + log2f(s[r]);
      }
    }
  }
}

template <int HDM>
__global__ void __launch_bounds__(ATT_THREADS) mhsa_bwd_rb_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                  const float* __restrict__ v, const float* __restrict__ o,
                                                                  const float* __restrict__ lse, const float* __restrict__ dout, int L,
                                                                  int E, int nh, float scale, float* __restrict__ dq,
                                                                  float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm_att[];
  float* sA = sm_att;                          // phase A: K      phase B: Q
  float* sB = sm_att + (size_t)L * E;          // phase A: V      phase B: dO
  float2* sLD = reinterpret_cast<float2*>(sB + (size_t)L * E);      // [nh][L] (log-sum-exp, D = rowsum(dO * O))
  const int64_t b = blockIdx.x;
  const size_t base = (size_t)b * L * E;
  const float scale_log2e = scale * 1.4426950408889634f;
  stage_rows(sA, k + base, L * E);
  stage_rows(sB, v + base, L * E);
  for (int it = threadIdx.x; it < L * nh; it += blockDim.x) {
    const int h = it / L, l = it - h * L;
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < HDM; ++i) d = fmaf(dout[base + (size_t)l * E + h * HDM + i], o[base + (size_t)l * E + h * HDM + i], d);
    sLD[it] = make_float2(lse[(size_t)b * nh * L + it], d);
  }
  __syncthreads();
  const int G = (L + ATT_RB - 1) / ATT_RB;
  // ---- phase A: item = (head, 4 query rows) -> dQ
  for (int item = threadIdx.x; item < nh * G; item += blockDim.x) {
    const int h = item / G, g = item - h * G;
    float qv[ATT_RB][HDM], dov[ATT_RB][HDM], acc[ATT_RB][HDM], ls[ATT_RB], Dl[ATT_RB];
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int l = g + r * G;
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        qv[r][i] = l < L ? q[base + (size_t)l * E + h * HDM + i] * scale_log2e : 0.f;
        dov[r][i] = l < L ? dout[base + (size_t)l * E + h * HDM + i] : 0.f;
        acc[r][i] = 0.f;
      }
      const float2 t = l < L ? sLD[h * L + l] : make_float2(0.f, 0.f);
      ls[r] = t.x;
      Dl[r] = t.y;
    }
    const float* kr = sA + h * HDM;
    const float* vr = sB + h * HDM;
    for (int j = 0; j < L; ++j, kr += E, vr += E) {
      float kk[HDM], vv[HDM];
      load_head<HDM, true>(kk, kr, HDM);
      load_head<HDM, true>(vv, vr, HDM);
#pragma unroll
      for (int r = 0; r < ATT_RB; ++r) {
        const float d = dot_p<HDM>(qv[r], kk), dp = dot_p<HDM>(dov[r], vv);
        const float ds = fast_exp2(d - ls[r]) * (dp - Dl[r]);
        axpy_p<HDM>(ds, kk, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int l = g + r * G;
      if (l < L) {
#pragma unroll
        for (int i = 0; i < HDM; ++i) dq[base + (size_t)l * E + h * HDM + i] = acc[r][i] * scale;
      }
    }
  }
  __syncthreads();
  // ---- phase B: item = (head, 4 key rows) -> dK, dV
  stage_rows(sA, q + base, L * E);
  stage_rows(sB, dout + base, L * E);
  __syncthreads();
  for (int item = threadIdx.x; item < nh * G; item += blockDim.x) {
    const int h = item / G, g = item - h * G;
    float kv[ATT_RB][HDM], vv[ATT_RB][HDM], dkv[ATT_RB][HDM], dvv[ATT_RB][HDM];
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int j = g + r * G;
#pragma unroll
      for (int i = 0; i < HDM; ++i) {
        kv[r][i] = j < L ? k[base + (size_t)j * E + h * HDM + i] * scale_log2e : 0.f;
        vv[r][i] = j < L ? v[base + (size_t)j * E + h * HDM + i] : 0.f;
        dkv[r][i] = 0.f;
        dvv[r][i] = 0.f;
      }
    }
    const float* qr = sA + h * HDM;
    const float* dor = sB + h * HDM;
    const float2* ldr = sLD + h * L;
    for (int l = 0; l < L; ++l, qr += E, dor += E) {
      float qq[HDM], dd[HDM];
      load_head<HDM, true>(qq, qr, HDM);
      load_head<HDM, true>(dd, dor, HDM);
      const float2 t = ldr[l];
#pragma unroll
      for (int r = 0; r < ATT_RB; ++r) {
        const float d = dot_p<HDM>(qq, kv[r]), dp = dot_p<HDM>(dd, vv[r]);
        const float p = fast_exp2(d - t.x);
        const float ds = p * (dp - t.y);
        axpy_p<HDM>(p, dd, dvv[r]);
        axpy_p<HDM>(ds, qq, dkv[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ATT_RB; ++r) {
      const int j = g + r * G;
      if (j < L) {
#pragma unroll
        for (int i = 0; i < HDM; ++i) {
          dk[base + (size_t)j * E + h * HDM + i] = dkv[r][i] * scale;
          dv[base + (size_t)j * E + h * HDM + i] = dvv[r][i];
        }
      }
    }
  }
}

int g_mhsa_row_blocked = 1;     // 1: row-blocked kernels whenever head_dim is exactly 2 or 4 (default); 0: one row per thread
extern "C" void xdfm_mhsa_set_row_blocked(int v) { g_mhsa_row_blocked = v ? 1 : 0; }

static int mhsa_check(int64_t B, int L, int E, int nh, size_t smem_floats, const char* what) {
  XDFM_CHECK_ARG(B >= 0 && L >= 1 && E >= 1 && nh >= 1 && E % nh == 0, "%s: bad shape B=%lld L=%d E=%d heads=%d", what, (long long)B, L,
                 E, nh);
  if (E / nh > 32) {
    xdfm_set_error("%s: head_dim %d > 32 is not supported by the fused attention kernel", what, E / nh);
    return XDFM_ERR_UNSUPPORTED;
  }
  if (smem_floats * 4 > 220 * 1024) {
    xdfm_set_error("%s: L*E = %d*%d does not fit the shared memory of one SM", what, L, E);
    return XDFM_ERR_UNSUPPORTED;
  }
  return XDFM_OK;
}

static int make_drop_cfg(float p, const void* seed_dev, DropCfg* cfg, const char* what) {
  XDFM_CHECK_ARG(p >= 0.f && p < 1.f, "%s: dropout p=%g must be in [0, 1)", what, (double)p);
  XDFM_CHECK_ARG(p == 0.f || seed_dev != nullptr, "%s: dropout needs a device seed", what);
  cfg->seed = (const unsigned long long*)seed_dev;
  cfg->threshold = (unsigned int)std::min(4294967295.0, (double)p * 4294967296.0);
  cfg->inv_keep = 1.f / (1.f - p);
  return XDFM_OK;
}

static int mhsa_fwd_impl(const float* q, const float* k, const float* v, int64_t B, int L, int E, int heads, float p, const void* seed_dev,
                         float* o, float* lse, void* stream) {
  DropCfg drop;
  int rc0 = make_drop_cfg(p, seed_dev, &drop, "mhsa_fwd");
  if (rc0) return rc0;
  const bool use_drop = p > 0.f;
  const size_t smem = (size_t)2 * L * E;
  int rc = mhsa_check(B, L, E, heads, smem, "mhsa_fwd");
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  const int hd = E / heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)hd);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_FWD_(HDM, EX, DR)                                                                                            \
  do {                                                                                                                      \
    XDFM_CUDA(cudaFuncSetAttribute(mhsa_fwd_kernel<HDM, EX, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * 4))); \
    mhsa_fwd_kernel<HDM, EX, DR><<<(unsigned)B, ATT_THREADS, smem * 4, st>>>(q, k, v, L, E, heads, hd, scale_log2e, o, lse, drop); \
  } while (0)
#define LAUNCH_FWD(HDM)                                                                                                     \
  do {                                                                                                                      \
    if (use_drop) { if (hd == HDM) LAUNCH_FWD_(HDM, true, true); else LAUNCH_FWD_(HDM, false, true); }                      \
    else { if (hd == HDM) LAUNCH_FWD_(HDM, true, false); else LAUNCH_FWD_(HDM, false, false); }                             \
  } while (0)
#define LAUNCH_FWD_RB(HDM)                                                                                                  \
  do {                                                                                                                      \
    XDFM_CUDA(cudaFuncSetAttribute(mhsa_fwd_rb_kernel<HDM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * 4)));  \
    mhsa_fwd_rb_kernel<HDM><<<(unsigned)B, ATT_THREADS, smem * 4, st>>>(q, k, v, L, E, heads, scale_log2e, o, lse);          \
  } while (0)
  // wider heads need 4 x 4 x head_dim registers per thread in the backward (8: 177, 16: spills): one row per thread there
  // (dropout runs on the one-row-per-thread kernels: the hash per (query, key) pair dominates there anyway)
  const bool rb = !use_drop && g_mhsa_row_blocked && ((hd == 2 && E % 2 == 0) || (hd == 4 && E % 4 == 0));
  if (rb && hd == 2) LAUNCH_FWD_RB(2); else if (rb && hd == 4) LAUNCH_FWD_RB(4);
  else if (hd <= 2) LAUNCH_FWD(2); else if (hd <= 4) LAUNCH_FWD(4); else if (hd <= 8) LAUNCH_FWD(8); else if (hd <= 16) LAUNCH_FWD(16);
  else LAUNCH_FWD(32);
#undef LAUNCH_FWD_RB
#undef LAUNCH_FWD
#undef LAUNCH_FWD_
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_mhsa_fwd(const float* q, const float* k, const float* v, int64_t B, int L, int E, int heads, float* o, float* lse,
                             void* stream) {
  return mhsa_fwd_impl(q, k, v, B, L, E, heads, 0.f, nullptr, o, lse, stream);
}

extern "C" int xdfm_mhsa_fwd_dropout(const float* q, const float* k, const float* v, int64_t B, int L, int E, int heads, float p,
                                     const void* seed_dev, float* o, float* lse, void* stream) {
  return mhsa_fwd_impl(q, k, v, B, L, E, heads, p, seed_dev, o, lse, stream);
}

static int mhsa_bwd_impl(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dout,
                         int64_t B, int L, int E, int heads, float p, const void* seed_dev, float* dq, float* dk, float* dv, void* stream) {
  DropCfg drop;
  int rc0 = make_drop_cfg(p, seed_dev, &drop, "mhsa_bwd");
  if (rc0) return rc0;
  const bool use_drop = p > 0.f;
  const size_t smem = (size_t)2 * L * E + (size_t)2 * L * heads;
  int rc = mhsa_check(B, L, E, heads, smem, "mhsa_bwd");
  if (rc) return rc;
  if (B == 0) return XDFM_OK;
  const int hd = E / heads;
  const float scale = 1.f / sqrtf((float)hd);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_BWD_(HDM, EX, DR)                                                                                            \
  do {                                                                                                                      \
    XDFM_CUDA(cudaFuncSetAttribute(mhsa_bwd_kernel<HDM, EX, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * 4))); \
    mhsa_bwd_kernel<HDM, EX, DR><<<(unsigned)B, ATT_THREADS, smem * 4, st>>>(q, k, v, o, lse, dout, L, E, heads, hd, scale, dq, dk, dv, drop); \
  } while (0)
#define LAUNCH_BWD(HDM)                                                                                                     \
  do {                                                                                                                      \
    if (use_drop) { if (hd == HDM) LAUNCH_BWD_(HDM, true, true); else LAUNCH_BWD_(HDM, false, true); }                      \
    else { if (hd == HDM) LAUNCH_BWD_(HDM, true, false); else LAUNCH_BWD_(HDM, false, false); }                             \
  } while (0)
#define LAUNCH_BWD_RB(HDM)                                                                                                  \
  do {                                                                                                                      \
    XDFM_CUDA(cudaFuncSetAttribute(mhsa_bwd_rb_kernel<HDM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * 4)));  \
    mhsa_bwd_rb_kernel<HDM><<<(unsigned)B, ATT_THREADS, smem * 4, st>>>(q, k, v, o, lse, dout, L, E, heads, scale, dq, dk, dv); \
  } while (0)
  const bool rb = !use_drop && g_mhsa_row_blocked && ((hd == 2 && E % 2 == 0) || (hd == 4 && E % 4 == 0));
  if (rb && hd == 2) LAUNCH_BWD_RB(2); else if (rb && hd == 4) LAUNCH_BWD_RB(4);
  else if (hd <= 2) LAUNCH_BWD(2); else if (hd <= 4) LAUNCH_BWD(4); else if (hd <= 8) LAUNCH_BWD(8); else if (hd <= 16) LAUNCH_BWD(16);
  else LAUNCH_BWD(32);
#undef LAUNCH_BWD_RB
#undef LAUNCH_BWD
#undef LAUNCH_BWD_
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_mhsa_bwd(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dout,
                             int64_t B, int L, int E, int heads, float* dq, float* dk, float* dv, void* stream) {
  return mhsa_bwd_impl(q, k, v, o, lse, dout, B, L, E, heads, 0.f, nullptr, dq, dk, dv, stream);
}

extern "C" int xdfm_mhsa_bwd_dropout(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dout,
                                     int64_t B, int L, int E, int heads, float p, const void* seed_dev, float* dq, float* dk, float* dv,
                                     void* stream) {
  return mhsa_bwd_impl(q, k, v, o, lse, dout, B, L, E, heads, p, seed_dev, dq, dk, dv, stream);
}

// the keep mask the dropout kernels recompute, materialised (tests / debugging): mask [B, heads, L, L] uint8
__global__ void mhsa_dropout_mask_kernel(int64_t total, int L, int nh, DropCfg drop, unsigned char* __restrict__ mask) {
  const unsigned long long seed = *drop.seed;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % L);
    mask[i] = drop_keep(seed, (unsigned long long)(i / L), j, drop.threshold) ? 1 : 0;
  }
}

extern "C" int xdfm_mhsa_dropout_mask(int64_t B, int L, int heads, float p, const void* seed_dev, unsigned char* mask, void* stream) {
  DropCfg drop;
  int rc = make_drop_cfg(p, seed_dev, &drop, "mhsa_dropout_mask");
  if (rc) return rc;
  XDFM_CHECK_ARG(seed_dev != nullptr, "mhsa_dropout_mask: needs a device seed");
  const int64_t total = B * heads * (int64_t)L * L;
  if (total == 0) return XDFM_OK;
  int blocks = (int)std::min<int64_t>((int64_t)xdfm_num_sms() * 8, ceil_div64(total, 256));
  mhsa_dropout_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(total, L, heads, drop, mask);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// y = LayerNorm(a + r) * gamma + beta over the last dim E (r, gamma/beta optional); one thread per row
// ------------------------------------------------------------------------------------------------
#define LN_EMAX 64

template <int EM>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const float* __restrict__ a, const float* __restrict__ r,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
                                                         int E, float eps, int normalize, float* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd) {
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    float x[EM];
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < EM; ++i) {
      float t = 0.f;
      if (i < E) {
        t = a[row * E + i];
        if (r != nullptr) t += r[row * E + i];
      }
      x[i] = t;
      mu += t;
    }
    if (!normalize) {
#pragma unroll
      for (int i = 0; i < EM; ++i)
        if (i < E) y[row * E + i] = x[i];
      continue;
    }
    mu /= (float)E;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < EM; ++i)
      if (i < E) { const float d = x[i] - mu; var = fmaf(d, d, var); }
    const float rs = rsqrtf(var / (float)E + eps);
#pragma unroll
    for (int i = 0; i < EM; ++i)
      if (i < E) y[row * E + i] = (x[i] - mu) * rs * gamma[i] + beta[i];
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx (= d a = d r) and per-block partial sums of dgamma / dbeta: partial [gridDim.x, 2E]
template <int EM>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                         const float* __restrict__ r, const float* __restrict__ gamma,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd, int64_t rows, int E,
                                                         float* __restrict__ dx, float* __restrict__ partial) {
  __shared__ float red[8][2 * EM];
  float dg[EM], db[EM];
#pragma unroll
  for (int i = 0; i < EM; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float xh[EM], g[EM];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < EM; ++i) {
      xh[i] = 0.f;
      g[i] = 0.f;
      if (i < E) {
        float t = a[row * E + i];
        if (r != nullptr) t += r[row * E + i];
        xh[i] = (t - mu) * rs;
        const float d = dy[row * E + i];
        g[i] = d * gamma[i];
        dg[i] = fmaf(d, xh[i], dg[i]);
        db[i] += d;
        m1 += g[i];
        m2 = fmaf(g[i], xh[i], m2);
      }
    }
    m1 /= (float)E;
    m2 /= (float)E;
#pragma unroll
    for (int i = 0; i < EM; ++i)
      if (i < E) dx[row * E + i] = rs * (g[i] - m1 - xh[i] * m2);
  }
  // block reduction in a fixed order (deterministic): warp shuffle tree, then the 8 warps in order
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < EM; ++i) {
    if (i < E) {
      const float s1 = warp_sum(dg[i]), s2 = warp_sum(db[i]);
      if (lane == 0) { red[w][i] = s1; red[w][E + i] = s2; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
    float t = 0.f;
    for (int ww = 0; ww < 8; ++ww) t += red[ww][i];
    partial[(size_t)blockIdx.x * 2 * E + i] = t;
  }
}

extern "C" int xdfm_add_ln_fwd(const float* a, const float* r, const float* gamma, const float* beta, int64_t rows, int E, float eps,
                               int normalize, float* y, float* mean, float* rstd, void* stream) {
  XDFM_CHECK_ARG(E >= 1 && E <= LN_EMAX, "add_ln_fwd: E=%d unsupported (1..%d)", E, LN_EMAX);
  XDFM_CHECK_ARG(!normalize || (gamma && beta && mean && rstd), "add_ln_fwd: gamma/beta/mean/rstd required");
  if (rows == 0) return XDFM_OK;
  int blocks = (int)min((int64_t)xdfm_num_sms() * 8, ceil_div64(rows, 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (E <= 8) add_ln_fwd_kernel<8><<<blocks, 256, 0, st>>>(a, r, gamma, beta, rows, E, eps, normalize, y, mean, rstd);
  else if (E <= 16) add_ln_fwd_kernel<16><<<blocks, 256, 0, st>>>(a, r, gamma, beta, rows, E, eps, normalize, y, mean, rstd);
  else if (E <= 32) add_ln_fwd_kernel<32><<<blocks, 256, 0, st>>>(a, r, gamma, beta, rows, E, eps, normalize, y, mean, rstd);
  else add_ln_fwd_kernel<64><<<blocks, 256, 0, st>>>(a, r, gamma, beta, rows, E, eps, normalize, y, mean, rstd);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_add_ln_bwd_blocks(int64_t rows) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)xdfm_num_sms() * 4, ceil_div64(rows, 256)));
}

extern "C" int xdfm_add_ln_bwd(const float* dy, const float* a, const float* r, const float* gamma, const float* mean, const float* rstd,
                               int64_t rows, int E, float* dx, float* partial, void* stream) {
  XDFM_CHECK_ARG(E >= 1 && E <= LN_EMAX, "add_ln_bwd: E=%d unsupported (1..%d)", E, LN_EMAX);
  if (rows == 0) return XDFM_OK;
  int blocks = xdfm_add_ln_bwd_blocks(rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (E <= 8) add_ln_bwd_kernel<8><<<blocks, 256, 0, st>>>(dy, a, r, gamma, mean, rstd, rows, E, dx, partial);
  else if (E <= 16) add_ln_bwd_kernel<16><<<blocks, 256, 0, st>>>(dy, a, r, gamma, mean, rstd, rows, E, dx, partial);
  else if (E <= 32) add_ln_bwd_kernel<32><<<blocks, 256, 0, st>>>(dy, a, r, gamma, mean, rstd, rows, E, dx, partial);
  else add_ln_bwd_kernel<64><<<blocks, 256, 0, st>>>(dy, a, r, gamma, mean, rstd, rows, E, dx, partial);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

// ------------------------------------------------------------------------------------------------
// attention pooling: a = softmax_L(score), out[b, :] = sum_l a[b,l] * x[b,l,:]; one CTA per sample
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

__global__ void __launch_bounds__(256) attn_pool_fwd_kernel(const float* __restrict__ score, const float* __restrict__ x, int L, int E,
                                                            float* __restrict__ attn, float* __restrict__ out) {
  extern __shared__ float sm_pool[];      // a [L] + partial [256]
  float* sA = sm_pool;
  float* sP = sm_pool + L;
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  float mx = -CUDART_INF_F;
  for (int l = threadIdx.x; l < L; l += blockDim.x) mx = fmaxf(mx, score[b * L + l]);
  mx = block_reduce(mx, red, true);
  float s = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float e = __expf(score[b * L + l] - mx);
    sA[l] = e;
    s += e;
  }
  s = block_reduce(s, red, false);
  const float inv = 1.f / s;
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float av = sA[l] * inv;
    sA[l] = av;
    attn[b * L + l] = av;
  }
  __syncthreads();
  // thread t: channel e = t % EP, row group = t / EP
  int EP = 1;
  while (EP < E) EP <<= 1;
  const int groups = blockDim.x / EP;
  const int e = threadIdx.x % EP, gq = threadIdx.x / EP;
  float acc = 0.f;
  if (e < E && gq < groups)
    for (int l = gq; l < L; l += groups) acc = fmaf(sA[l], x[((size_t)b * L + l) * E + e], acc);
  sP[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < E) {
    float t = 0.f;
    for (int g2 = 0; g2 < groups; ++g2) t += sP[g2 * EP + threadIdx.x];
    out[b * E + threadIdx.x] = t;
  }
}

// dscore[b,l] = a_l * (da_l - sum_l' a_l' da_l'),  da_l = dout[b,:] . x[b,l,:];   dx[b,l,:] = a_l * dout[b,:]
__global__ void __launch_bounds__(256) attn_pool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ attn,
                                                            const float* __restrict__ x, int L, int E, float* __restrict__ dscore,
                                                            float* __restrict__ dx) {
  extern __shared__ float sm_pool[];      // dout [E] + da [L]
  float* sG = sm_pool;
  float* sDa = sm_pool + E;
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < E; i += blockDim.x) sG[i] = dout[b * E + i];
  __syncthreads();
  float dot = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float* xr = x + ((size_t)b * L + l) * E;
    float da = 0.f;
    for (int i = 0; i < E; ++i) da = fmaf(sG[i], xr[i], da);
    sDa[l] = da;
    const float av = attn[b * L + l];
    dot = fmaf(av, da, dot);
    float* dr = dx + ((size_t)b * L + l) * E;
    for (int i = 0; i < E; ++i) dr[i] = av * sG[i];
  }
  dot = block_reduce(dot, red, false);
  for (int l = threadIdx.x; l < L; l += blockDim.x) dscore[b * L + l] = attn[b * L + l] * (sDa[l] - dot);
}

extern "C" int xdfm_attn_pool_fwd(const float* score, const float* x, int64_t B, int L, int E, float* attn, float* out, void* stream) {
  XDFM_CHECK_ARG(L >= 1 && E >= 1 && E <= 256, "attn_pool_fwd: L=%d E=%d unsupported", L, E);
  if (B == 0) return XDFM_OK;
  size_t smem = ((size_t)L + 256) * 4;
  XDFM_CHECK_ARG(smem <= 48 * 1024, "attn_pool_fwd: L=%d too long", L);
  attn_pool_fwd_kernel<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(score, x, L, E, attn, out);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}

extern "C" int xdfm_attn_pool_bwd(const float* dout, const float* attn, const float* x, int64_t B, int L, int E, float* dscore, float* dx,
                                  void* stream) {
  XDFM_CHECK_ARG(L >= 1 && E >= 1 && E <= 256, "attn_pool_bwd: L=%d E=%d unsupported", L, E);
  if (B == 0) return XDFM_OK;
  size_t smem = ((size_t)L + E) * 4;
  XDFM_CHECK_ARG(smem <= 48 * 1024, "attn_pool_bwd: L=%d too long", L);
  attn_pool_bwd_kernel<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(dout, attn, x, L, E, dscore, dx);
  XDFM_LAUNCH_CHECK();
  return XDFM_OK;
}
