/* libxdfm_sm100a.so -- C ABI of the B200-native xDeepFM hot path.
 *
 * The reference (Syclus123/xDeepFM-pytorch) is pure Python/PyTorch and has NO FFI boundary of its own; its
 * "operator API" is the deepctr Python class surface.  This header is therefore a NEW boundary: every entry point
 * names the reference code (file:line under /root/reference) whose arithmetic it replaces.  The Python host side
 * (xdeepfm-pytorch_b200/deepctr) binds these symbols with ctypes and keeps the reference's
 * class/ctor/state_dict surface.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless marked "host".
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on that stream (no sync, no
 *     allocation; workspaces are caller-provided, sized by the matching *_workspace_bytes()).
 *   - return value 0 = ok; non-zero = error, message via xdfm_last_error().
 *   - row-major fp32 unless stated.  B = batch, m = sparse fields, D = embedding dim, nd = dense features.
 */
#ifndef XDFM_H_
#define XDFM_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XDFM_MAX_FIELDS 64
#define XDFM_MAX_DENSE 256

#define XDFM_ACT_NONE 0
#define XDFM_ACT_RELU 1
#define XDFM_ACT_TANH 2
#define XDFM_ACT_SIGMOID 3

#define XDFM_OPT_SGD 0
#define XDFM_OPT_ADAM 1
#define XDFM_OPT_ADAGRAD 2
#define XDFM_OPT_RMSPROP 3

/* optimizer hyper-parameters (torch.optim defaults are applied by the Python side: basemodel.py:447-461) */
typedef struct {
  int32_t kind;     /* XDFM_OPT_* */
  float lr;
  float beta1;      /* adam */
  float beta2;      /* adam */
  float eps;
  float alpha;      /* rmsprop */
  float lr_decay;   /* adagrad */
  float l2;         /* L2 regulariser strength of the parameter group: loss term l2*sum(w^2) -> grad 2*l2*w */
} xdfm_opt_cfg;

const char* xdfm_last_error(void);
int xdfm_version(void);
/* compute capability of the current device * 10 (100 on B200); <0 on error */
int xdfm_device_cc(void);
/* kernels of this library launched so far by this process */
long long xdfm_launch_count(void);

/* ---- input split: replaces X[:, i:i+1].long() / X[:, a:b] column slicing (deepctr/models/basemodel.py:368-370, 377-378).
 * X [B, ncol] fp32 (ids stored as floats); sparse_cols/dense_cols are HOST arrays of column indices.
 * ids [B, m] int32 (truncation toward zero == .long()), dense [B, nd] fp32. */
int xdfm_split_input(const float* X, int64_t B, int ncol, const int32_t* sparse_cols, int m, const int32_t* dense_cols, int nd,
                     int32_t* ids, float* dense, void* stream);

/* ---- fused multi-table embedding gather + first-order term.
 * replaces input_from_feature_columns (basemodel.py:354-380), torch.cat(.., dim=1) (xdeepfm.py:86) and
 * Linear.forward (basemodel.py:63-92).
 * tables / lin_tables: HOST arrays (length m, one entry per FEATURE) of device pointers to [V_f, D] / [V_f, 1] fp32.
 * vocab: HOST array of V_f (ids are clamped into range).  out_emb [B, m, D]; out_lin [B] = sum_f lin_f[id] + dense.dense_w.
 * out_emb or out_lin may be NULL to skip that half. */
int xdfm_embed_gather(const float* const* tables, const float* const* lin_tables, const int32_t* vocab, const int32_t* ids, int64_t B,
                      int m, int D, float* out_emb, const float* dense, int nd, const float* dense_w, float* out_lin, void* stream);

/* ---- deterministic sorted segmented scatter-add (backward of the gather; replaces aten::embedding_dense_backward).
 * step 1: keys = row_offset[f] + ids[b,f] -> radix sort -> run-length segments.
 *   row_offset: HOST array (length m) = first global row of feature f's table (features sharing a table share it).
 *   outputs: uniq_keys [n_keys] (first *num_segments valid), seg_offsets [n_keys+1], sorted_pos [n_keys] (index b*m+f),
 *   num_segments [1] (device). n_keys = B*m. */
int64_t xdfm_embed_bwd_workspace_bytes(int64_t n_keys);
int xdfm_embed_bwd_segments(const int32_t* ids, int64_t B, int m, const int64_t* row_offset, const int32_t* vocab, int64_t total_rows,
                            void* workspace, int64_t workspace_bytes, uint32_t* uniq_keys, int32_t* seg_offsets, int32_t* sorted_pos,
                            int32_t* num_segments, void* stream);
/* step 2: gsum[s, :] = sum over the segment's entries of demb[pos, :] ([B*m, D] view of d(out_emb)); gsum_lin[s] = sum of
 * dlin[pos / m].  Summation order is a fixed function of the segment length -> bit-reproducible. demb or dlin may be NULL. */
int xdfm_embed_bwd_reduce(const float* demb, const float* dlin, const int32_t* sorted_pos, const int32_t* seg_offsets,
                          const int32_t* num_segments, int64_t n_keys, int m, int D, float* gsum, float* gsum_lin, void* stream);
/* optional: scatter the segment sums into dense per-table gradients (what nn.Embedding(sparse=False) would produce).
 * grad_tables: HOST array (length T) of device pointers [rows_t, width]; table_row_offset: HOST array length T+1. */
int xdfm_embed_bwd_scatter_dense(float* const* grad_tables, const int64_t* table_row_offset, int T, int width, const uint32_t* uniq_keys,
                                 const float* gsum, const int32_t* num_segments, int64_t max_segments, void* stream);

/* ---- CIN layer, fp32 CUDA-core path (deepctr/layers/interaction.py:207-248).
 * x0 [B,m,D]; xk = layer input: element (b,i,d) at xk[b*xk_bstride + i*D + d], i < Hp; W [H, Hp*m]; bias [H].
 * y [B,H,D] = act(W . (xk (x) x0) + bias).  Channels [direct_begin, H) are "direct": their sum over D goes to
 * pooled[b, col_off + h - direct_begin] (pooled is [B, fm_total]) and/or the un-pooled values to maps [B, fm_total, D]. */
int xdfm_cin_fwd_f32(const float* x0, const float* xk, int64_t xk_bstride, const float* W, const float* bias, int64_t B, int m, int Hp,
                     int H, int D, int act, float* y, int direct_begin, float* pooled, float* maps, int fm_total, int col_off,
                     void* stream);
/* dy = act'(y) * (dpooled / dmaps on direct channels + dnext [B, n_next, D] on channels [0, n_next)) */
int xdfm_cin_dy(const float* y, int64_t B, int H, int D, int act, int direct_begin, const float* dpooled, const float* dmaps,
                int fm_total, int col_off, const float* dnext, int n_next, float* dy, void* stream);
/* dW [H, Hp*m], db [H] (overwritten); dxk [B,Hp,D] (overwritten); dx0 [B,m,D] (accumulated +=). */
int64_t xdfm_cin_bwd_f32_workspace_bytes(int64_t B, int m, int Hp, int H, int D);
int xdfm_cin_bwd_f32(const float* x0, const float* xk, int64_t xk_bstride, const float* W, const float* dy, int64_t B, int m, int Hp,
                     int H, int D, float* dW, float* db, float* dxk, float* dx0, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- dense helpers (DNN: deepctr/layers/core.py:120-134; heads: xdeepfm.py:56,73,88-105; core.py:154-160) */
int64_t xdfm_gemm_workspace_bytes(int M, int N, int K);
/* C[M,N] = act(op(A)[M,K] op(B)[K,N] + bias[N]) (+= C if accumulate) */
int xdfm_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                  const float* bias, int act, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);
int xdfm_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, void* stream);
int64_t xdfm_wcolsum_workspace_bytes(int K);
/* out[k] (+)= sum_b s[b] * X[b*ld + k]; s may be NULL (plain column sum); deterministic */
int xdfm_wcolsum(const float* X, int64_t B, int K, int ld, const float* s, float* out, int accumulate, void* workspace,
                 int64_t workspace_bytes, void* stream);
/* y_pred[b] = sigmoid(lin[b] + dnn_out[b,:].w_dnn + cin_out[b,:].w_cin + bias) (binary) or the raw logit */
int xdfm_head_fwd(const float* lin, const float* cin_out, const float* w_cin, int fm, const float* dnn_out, const float* w_dnn, int hd,
                  const float* bias, int64_t B, int binary, float* y_pred, void* stream);
int xdfm_head_bwd(const float* dy_pred, const float* y_pred, int64_t B, int binary, const float* w_cin, int fm, const float* w_dnn,
                  int hd, float* dlogit, float* d_cin_out, float* d_dnn_out, void* stream);
/* Backward of the head in one pass (autograd of xdeepfm.py:88-105): dlogit, d_cin_out, d_dnn_out as xdfm_head_bwd, plus
 * d_w [fm + hd + 1] = (d_w_cin = sum_b dlogit[b] * cin_out[b,:] | d_w_dnn | d_bias = sum_b dlogit[b]); fixed summation order.
 * cin_out / dnn_out NULL = that branch is absent (fm / hd count as 0). */
int64_t xdfm_head_bwd_fused_workspace_bytes(int64_t B, int fm, int hd);
int xdfm_head_bwd_fused(const float* dy_pred, const float* y_pred, int64_t B, int binary, const float* cin_out, const float* w_cin, int fm,
                        const float* dnn_out, const float* w_dnn, int hd, float* dlogit, float* d_cin_out, float* d_dnn_out, float* d_w,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* F.binary_cross_entropy(y_pred, y, reduction='sum') (basemodel.py:254): *loss_sum += loss; dy_pred = scale * dL/dy_pred */
int xdfm_bce_sum(const float* y_pred, const float* labels, int64_t B, float scale, float* per_sample, float* dy_pred,
                 double* loss_sum, void* stream);

/* ---- fused optimizer + L2 regulariser (basemodel.py:262, 412-428, 447-461).
 * opt_dev: 8 floats of device state per optimizer instance (zero-initialised); xdfm_opt_tick advances the step counter
 * and recomputes the bias corrections ON DEVICE (CUDA-graph friendly). */
int xdfm_opt_tick(float* opt_dev, const xdfm_opt_cfg* cfg, void* stream);
/* all dense parameters as one flat buffer; l2vec = per-element L2 strength (may be NULL); *reg_out += sum(l2*w^2) */
int xdfm_flat_opt(const xdfm_opt_cfg* cfg, const float* opt_dev, int64_t n, float* w, const float* g, float* s1, float* s2,
                  const float* l2vec, float grad_scale, double* reg_out, void* stream);
/* embedding tables: w/s1/s2 HOST arrays (length T) of device pointers [rows_t, width]; rows present in (uniq_keys, gsum)
 * get g = gsum*grad_scale + 2*l2*w; if dense_pass != 0 every other row gets g = 2*l2*w (the reference's dense semantics;
 * touched_bitmap = ceil(total_rows/32) uint32 scratch); *reg_out += sum(l2*w^2) over the rows visited. */
int xdfm_rows_opt(const xdfm_opt_cfg* cfg, const float* opt_dev, float* const* w, float* const* s1, float* const* s2,
                  const int64_t* table_row_offset, int T, int width, const uint32_t* uniq_keys, const float* gsum,
                  const int32_t* num_segments, int64_t max_segments, float grad_scale, uint32_t* touched_bitmap, int dense_pass,
                  double* reg_out, void* stream);

/* ---- lazy ("deferred catch-up") form of the reference's dense table semantics (basemodel.py:126, 412-428, 447-461): every row
 * of every table moves every step (g = 2*l2*w through the optimizer's moments); a row the batch does not touch evolves on its own,
 * so its update is postponed -- last[row] (int32, key space of the scatter-add) = step up to which it is current, hist = per-step
 * scalars (4 floats per step: adam step size, 1/sqrt(bias_correction2), adagrad clr, lr; slot = step - hist_base) -- and replayed in registers, bit-for-bit the arithmetic of the dense pass,
 * by whoever needs the row.  xdfm_opt_tick_hist = xdfm_opt_tick + history record.  xdfm_rows_catchup: rows in uniq_keys are
 * replayed to the completed-step count and written back (call before the forward lookup of a training step).
 * xdfm_rows_mark_current: last[row] = step for the rows just updated by xdfm_rows_opt(dense_pass = 0).  xdfm_rows_flush: all rows
 * (before predict / state_dict / at the end of an epoch).  reg_out accumulates l2*w^2 of every replayed (row, step). */
int xdfm_opt_tick_hist(float* opt_dev, const xdfm_opt_cfg* cfg, float* hist, int64_t hist_cap, int64_t hist_base, void* stream);
int xdfm_rows_catchup(const xdfm_opt_cfg* cfg, const float* opt_dev, const float* hist, int64_t hist_base, float* const* w,
                      float* const* s1, float* const* s2, int32_t* last, const int64_t* table_row_offset, int T, int width,
                      const uint32_t* uniq_keys, const int32_t* num_segments, int64_t max_segments, double* reg_out, void* stream);
int xdfm_rows_mark_current(int32_t* last, const uint32_t* uniq_keys, const int32_t* num_segments, int64_t max_segments,
                           const float* opt_dev, void* stream);
int xdfm_rows_flush(const xdfm_opt_cfg* cfg, const float* opt_dev, const float* hist, int64_t hist_base, float* const* w, float* const* s1,
                    float* const* s2, int32_t* last, const int64_t* table_row_offset, int T, int width, double* reg_out, void* stream);

/* lazy semantics for row-sharded tables: the READER replays.  ptrs_dev = DEVICE int64 [8*G], for every rank g the peer-mapped
 * addresses of its shard's (emb, lin, s1, s1_lin, s2, s2_lin, last, last_lin); rows behind the local step counter are replayed in registers
 * (nothing is written back: the owner catches its rows up when it applies the step). Other arguments as xdfm_embed_gather_sharded. */
int xdfm_embed_gather_sharded_lazy(const void* ptrs_dev, const int64_t* feat_base_dev, const int32_t* vocab, const int32_t* ids, int64_t B,
                                   int m, int D, int G, const xdfm_opt_cfg* cfg_emb, const xdfm_opt_cfg* cfg_lin, const float* opt_dev,
                                   const float* hist, int64_t hist_base, float* out_emb, const float* dense, int nd, const float* dense_w,
                                   float* out_lin, void* stream);

/* Row-sharded lookup through the batch's DISTINCT rows (replaces the nn.DataParallel replica lookup of basemodel.py:206-209, 354-380 at
 * G > 1): uniq_keys / seg_offsets / sorted_pos / num_segments are xdfm_shard_segments' outputs for this batch (the backward reuses them);
 * every distinct row crosses NVLink once into u_emb [nseg, D] / u_lin [nseg] (either may be NULL), replayed when stale (hist != NULL:
 * lazy tables; hist == NULL: tables are current, cfg_* / opt_dev ignored); inv [n] = segment of lookup q = b*m + f, n = B*m (NULL = skip).
 * xdfm_embed_expand_unique then writes out_emb [B, m, D] = u_emb[inv] and out_lin [B] = sum_f u_lin[inv[b, f]] + dense[b,:] . dense_w. */
int xdfm_embed_fetch_unique_sharded(const void* ptrs_dev, int G, uint32_t key_stride, int D, const uint32_t* uniq_keys,
                                    const int32_t* seg_offsets, const int32_t* sorted_pos, const int32_t* num_segments, int64_t n,
                                    const xdfm_opt_cfg* cfg_emb, const xdfm_opt_cfg* cfg_lin, const float* opt_dev, const float* hist,
                                    int64_t hist_base, float* u_emb, float* u_lin, int32_t* inv, void* stream);
int xdfm_embed_expand_unique(const float* u_emb, const float* u_lin, const int32_t* inv, int64_t B, int m, int D, float* out_emb,
                             const float* dense, int nd, const float* dense_w, float* out_lin, void* stream);
/* out [n] = u_lin[inv]: the first-order row of every lookup, un-summed (multi-value features are pooled per field first,
 * basemodel.py:63-92 with inputs.py:141-155) */
int xdfm_embed_expand_unique_lin_rows(const float* u_lin, const int32_t* inv, int64_t n, float* out, void* stream);

/* diagnostic: 0 = the lazy replay uses scalar arithmetic everywhere (default 1: packed fp32 pairs for 4-wide Adam pieces; same bits) */
int xdfm_set_replay_packed(int on);
/* diagnostic: 1 = first version of the dense-table streaming pass, 2 = unrolled / streaming-hint version (default) */
void xdfm_set_rows_opt_dense_version(int v);

/* ---- CIN layer on the tensor cores (bf16 operands, fp32 accumulate; deepctr/layers/interaction.py:218-246).
 * Activations use a ROW layout: one row per (sample, d), channels contiguous:
 *   x0t [B*D, mP] bf16 (mP = m rounded up to 8, zero padded; produced by xdfm_to_rows_bf16),
 *   xkt = layer input rows with pitch xk_pitch elements (layer k>0: the previous layer's yt, first Hp channels; layer 0: x0t),
 *   yt  [B*D, Hs] bf16 = act(W.(xk (x) x0) + bias), Hs = H rounded up to 8 (padding channels are written as zeros).
 * W fp32 [H, Hp*m] + bias [H] as the reference holds them; wprime = bf16 scratch of xdfm_cin_tc_wprime_elems() elements
 * (permuted/padded copy of W made by the call).  pooled / maps as in xdfm_cin_fwd_f32 (fp32, from the fp32 accumulators).
 * Supported: D in {8,16,32,64,128}, Hp <= 128, H <= 256; otherwise returns an error (use the fp32 path).
 * xdfm_cin_tc_set_cluster(c): thread-block cluster size (1, 2 or 4) used to multicast the weight stream (default 2). */
int xdfm_to_rows_bf16(const float* x, int64_t B, int C, int D, int CP, void* xt, void* stream);
int64_t xdfm_cin_tc_wprime_elems(int m, int Hp, int H, int D);
void xdfm_cin_tc_set_cluster(int c);
/* 1 (default): dual-producer forward kernel (row warps generate Z, then drain the accumulator) when D <= 32; 0: original kernel */
void xdfm_cin_tc_set_pair(int v);
int xdfm_cin_fwd_tc(const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, const float* bias, void* wprime, int64_t B,
                    int m, int Hp, int H, int D, int act, void* yt, int direct_begin, float* pooled, float* maps, int fm_total,
                    int col_off, void* stream);

/* row-layout helpers of the tensor-core path:
 * cin_dy_rows: dyt [B*D, Hs] bf16 = act'(yt) * (dpooled / dmaps on channels [direct_begin, H) + dnext [B*D, dnext_pitch] fp32
 *              (row layout) on channels [0, n_next));  from_rows_f32: [B*D, CP] fp32 rows -> [B, C, D] (optionally +=);
 * add_rows_f32: a[r, :C] += b[r, :C]. */
int xdfm_cin_dy_rows(const void* yt, int64_t B, int D, int H, int Hs, int direct_begin, const float* dpooled, const float* dmaps,
                     int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next, int act, void* dyt, void* stream);
/* fused: dyt [B*D, Hs] (row layout) AND dyT [H_pad, B*D] (channel-major, rows >= H zero) in one pass (= cin_dy_rows + rows_to_cols) */
int xdfm_cin_dy_rows_cols(const void* yt, int64_t B, int D, int H, int Hs, int H_pad, int direct_begin, const float* dpooled,
                          const float* dmaps, int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next, int act,
                          void* dyt, void* dyT, void* stream);
/* As xdfm_cin_dy_rows_cols; db != NULL also returns the layer's bias gradient db[h] = sum_r dY[r, h] (per-tile partial sums in
 * `workspace`, xdfm_cin_dy_db_workspace_bytes; fixed order) -- torch autograd's Conv1d bias gradient (interaction.py:224). */
int64_t xdfm_cin_dy_db_workspace_bytes(int64_t B, int D, int H_pad);
int xdfm_cin_dy_rows_cols_db(const void* yt, int64_t B, int D, int H, int Hs, int H_pad, int direct_begin, const float* dpooled,
                             const float* dmaps, int fm_total, int col_off, const float* dnext, int64_t dnext_pitch, int n_next, int act,
                             void* dyt, void* dyT, float* db, void* workspace, int64_t workspace_bytes, void* stream);
int xdfm_from_rows_f32(const float* xt, int64_t B, int C, int D, int CP, float* x, int accumulate, void* stream);
/* x [B, C, D] fp32 = extra[r, c] (NULL = none; row pitch extra_pitch) + the sum of the n_planes planes parts [n_planes, B*D, CP],
 * r = b * D + d: the CIN's dX^0 in the reference layout from the dX kernels' planes and layer 0's dXk */
int xdfm_cin_dx0_finish(const float* parts, int n_planes, const float* extra, int64_t extra_pitch, int64_t B, int C, int D, int CP, float* x,
                        void* stream);
int xdfm_add_rows_f32(float* a, int64_t pitch_a, const float* b, int64_t pitch_b, int64_t R, int C, void* stream);

/* CIN backward w.r.t. activations on the tensor cores: dyt [B*D, Hs] bf16 (act'(y) * upstream), x0t / xkt as in the forward,
 * wt = bf16 scratch [xdfm_cin_bwd_dx_tc_wt_elems]; dxk [B*D, HpQ] fp32 (overwritten, HpQ = Hp rounded up to 16),
 * dx0 [2, B*D, mP] fp32 (overwritten): this layer's dX^0 as two planes, one per half of the X^{k-1} channels; the layers' planes are
 * summed (and brought back to [B, m, D]) by xdfm_cin_dx0_finish.  (Replaces the dZ / einsum-backward part of torch autograd over
 * deepctr/layers/interaction.py:218-224.) */
int64_t xdfm_cin_bwd_dx_tc_wt_elems(int m, int Hp, int H, int D);
/* diagnostic bit mask for profiling experiments (0 = production; cfg2 layer shapes only): 1 skip the contraction FMAs, 2 TMEM loads,
 * 4 MMAs, 8 the weight stream, 16 tile outputs, 32 dY staging copies, 64 X^{k-1} loads; 256 = clock stamps follow one row warp */
void xdfm_cin_dx_set_debug(int v);
/* experiment switch of xdfm_cin_bwd_dw_tc: 0 = automatic fields per CTA (default), 1 = one field per CTA (deeper A ring in TMEM);
 * call before xdfm_cin_bwd_dw_tc_workspace_bytes / xdfm_cin_bwd_dw_tc of a step (both read it) */
void xdfm_cin_dw_set_jp(int v);
/* lane packing of xdfm_cin_bwd_dw_tc for narrow X^{k-1} (HpQ <= 64: 2 or 4 fields share the 128 TMEM lanes of one accumulator):
 * 1 = on (default), 0 = one field per accumulator (A/B tests); same calling rule as xdfm_cin_dw_set_jp */
void xdfm_cin_dw_set_pack(int enabled);
/* profiling only: device buffer of 2 * 32 * 16 int64 clock stamps written by CTA 0 of the following dX launches (NULL = off) */
void xdfm_cin_dx_set_trace(void* buf);
int xdfm_cin_bwd_dx_tc(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt, int64_t B, int m,
                       int Hp, int H, int D, float* dxk, float* dx0, void* stream);
/* As xdfm_cin_bwd_dx_tc, optionally writing the dY rows of the layer BELOW instead of dxk: when dy_prev != NULL,
 * dy_prev[r, i] = act'(xkt[r, i]) * dXk[r, i] for i < HpQ (bf16, row pitch dy_pitch >= HpQ, multiple of 8) and dxk may be NULL.  Valid
 * when the layer below feeds only this layer through those channels (split_half: its first Hp channels); act = XDFM_ACT_RELU / NONE.
 * The layer below then calls xdfm_cin_dy_rows_cols with dnext = NULL and dnext_pitch = -1 ("hidden half already in dyt"). */
int xdfm_cin_bwd_dx_tc_dy(const void* dyt, const void* x0t, const void* xkt, int64_t xk_pitch, const float* W, void* wt, int64_t B, int m,
                          int Hp, int H, int D, float* dxk, float* dx0, void* dy_prev, int64_t dy_pitch, int act, void* stream);

/* CIN weight gradient on the tensor cores.  Operands are CHANNEL-MAJOR bf16 copies (xdfm_rows_to_cols_bf16: rows [R, pitch]
 * -> [CP, R], rows >= C zero): dyT [H_pad, R], xkT [HpQ, R], x0T [mP, R] with H_pad/HpQ = H/Hp rounded up to 16, mP = m rounded
 * up to 8, R = B*D.  dW fp32 [H, Hp*m] (reference layout) and db [H] are overwritten; deterministic two-stage reduction. */
int xdfm_rows_to_cols_bf16(const void* src, int64_t pitch, int64_t R, int C, int CP, void* dst, void* stream);
int64_t xdfm_cin_bwd_dw_tc_workspace_bytes(int64_t B, int m, int Hp, int H, int D);
int xdfm_cin_bwd_dw_tc(const void* dyT, const void* xkT, const void* x0T, int64_t B, int m, int Hp, int H, int D, float* dW, float* db,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---- dense layers on the tensor cores (bf16 operands, fp32 accumulate): DNN.forward / autograd in bf16 precision
 * (deepctr/layers/core.py:120-134).  C[M,N] fp32 (row pitch ldc) = act(A[M,K] . B[N,K]^T + bias[N]); A / B are bf16 K-major with
 * row pitches lda / ldb (elements, multiples of 8) as produced by xdfm_cvt_bf16; workspace = xdfm_gemm_tc_workspace_bytes (split-K
 * partials, reduced in a fixed order).
 * xdfm_cvt_bf16: fp32 [R, C] (row pitch ld) -> bf16 [R, dst_pitch] or, transposed, [C, dst_pitch]; padding columns are zeros. */
int64_t xdfm_gemm_tc_workspace_bytes(int M, int N, int K);
int xdfm_gemm_tc(int M, int N, int K, const void* A, int64_t lda, const void* Bm, int64_t ldb, float* C, int ldc, const float* bias,
                 int act, void* workspace, int64_t workspace_bytes, void* stream);
int xdfm_cvt_bf16(const float* src, int R, int C, int64_t ld, int transpose, void* dst, int64_t dst_pitch, void* stream);
/* One pass for every bf16 operand of a dense layer's tensor-core GEMMs: g = src * act'(y) (y = NULL: g = src; fp32 [R, C], row pitch
 * ld for both), dst [R, dst_pitch] = bf16(g), dstT [C, dstT_pitch] = bf16(g)^T, colsum[c] = sum_r g[r, c] (fixed order; needs a
 * workspace of xdfm_cvt_bf16_both_workspace_bytes).  dst / dstT / colsum may each be NULL; pitches are multiples of 8, padding is zero.
 * Replaces torch autograd's ReLU-backward + the bias gradient sum of deepctr/layers/core.py:DNN together with the operand copies. */
int64_t xdfm_cvt_bf16_both_workspace_bytes(int R, int C);
int xdfm_cvt_bf16_both(const float* src, const float* y, int act, int R, int C, int64_t ld, void* dst, int64_t dst_pitch, void* dstT,
                       int64_t dstT_pitch, float* colsum, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- field self-attention block over the CIN feature maps (deepctr/layers/cin_attention.py).
 * q/k/v/o/dout [B, L, E] fp32 (E = heads * head_dim, head_dim <= 32); lse [B, heads, L] = log2-sum-exp2 of the scaled scores.
 * o = softmax(q k^T / sqrt(head_dim)) v per head (cin_attention.py:73-95, dropout p = 0); nothing of size L x L leaves the SM;
 * the backward recomputes the probabilities from lse. */
int xdfm_mhsa_fwd(const float* q, const float* k, const float* v, int64_t B, int L, int E, int heads, float* o, float* lse, void* stream);
int xdfm_mhsa_bwd(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dout, int64_t B, int L,
                  int E, int heads, float* dq, float* dk, float* dv, void* stream);
/* y = LayerNorm_E(a + r) * gamma + beta (cin_attention.py:305-311; r may be NULL; normalize = 0: y = a + r only).
 * mean / rstd [rows] are saved for the backward, which returns dx (= d a = d r) and per-block partial sums
 * partial [xdfm_add_ln_bwd_blocks(rows), 2E] of (dgamma, dbeta) to be column-summed (xdfm_wcolsum). E <= 64. */
/* diagnostic switch: 1 (default) = four query / key rows per thread and shared-memory load when head_dim is exactly 2 or 4;
 * 0 = one row per thread */
void xdfm_mhsa_set_row_blocked(int v);
/* the same attention core with dropout on the probabilities (nn.Dropout(attn_probs), deepctr/layers/cin_attention.py:54, 86):
 * keep mask = counter-based hash of (*seed_dev, sample, head, query, key) >= p * 2^32, recomputed by the backward (pass the same p
 * and a seed buffer holding the same value); kept probabilities are scaled by 1 / (1 - p).  seed_dev: device uint64 [1], read by the
 * kernel (so a CUDA-graph replay sees the value the host-side counter kernel left there).  xdfm_mhsa_dropout_mask materialises the
 * mask [B, heads, L, L] uint8 for tests. */
int xdfm_mhsa_fwd_dropout(const float* q, const float* k, const float* v, int64_t B, int L, int E, int heads, float p, const void* seed_dev,
                          float* o, float* lse, void* stream);
int xdfm_mhsa_bwd_dropout(const float* q, const float* k, const float* v, const float* o, const float* lse, const float* dout, int64_t B,
                          int L, int E, int heads, float p, const void* seed_dev, float* dq, float* dk, float* dv, void* stream);
int xdfm_mhsa_dropout_mask(int64_t B, int L, int heads, float p, const void* seed_dev, unsigned char* mask, void* stream);
int xdfm_add_ln_fwd(const float* a, const float* r, const float* gamma, const float* beta, int64_t rows, int E, float eps, int normalize,
                    float* y, float* mean, float* rstd, void* stream);
int xdfm_add_ln_bwd_blocks(int64_t rows);
int xdfm_add_ln_bwd(const float* dy, const float* a, const float* r, const float* gamma, const float* mean, const float* rstd,
                    int64_t rows, int E, float* dx, float* partial, void* stream);
/* attention pooling (cin_attention.py:138-142): attn = softmax over L of score [B, L]; out [B, E] = sum_l attn[b,l] x[b,l,:]. */
int xdfm_attn_pool_fwd(const float* score, const float* x, int64_t B, int L, int E, float* attn, float* out, void* stream);
int xdfm_attn_pool_bwd(const float* dout, const float* attn, const float* x, int64_t B, int L, int E, float* dscore, float* dx,
                       void* stream);

/* ---- narrow dense layers over very tall activations (K, N <= 32; up to three weight matrices sharing the input): the bias-free
 * E x E projections W_q / W_k / W_v / W_o of MultiHeadSelfAttention (deepctr/layers/cin_attention.py:49-61, 73-79, 95) and the
 * Linear(E,E)-Tanh-Linear(E,1) score MLP of AttentionPooling (cin_attention.py:114-118, 135-136), plus their autograd.
 * fwd:    y_q[R, N] = act(x[R, K] . w_q[N, K]^T + bias), q < nout (bias only with nout == 1; unused w / y pointers NULL).
 * bwd_dx: dx[R, K] = sum_q dy_q[R, N] . w_q[N, K].
 * bwd_dw: dw_q[N, K] = dy_q^T . x, db[N] = column sums of dy_0 (db may be NULL); deterministic two-stage reduction, workspace
 *         sized by xdfm_small_linear_bwd_dw_workspace_bytes (-1 = unsupported shape). */
int xdfm_small_linear_fwd(const float* x, const float* w0, const float* w1, const float* w2, const float* bias, int act, int64_t R,
                          int K, int N, int nout, float* y0, float* y1, float* y2, void* stream);
int xdfm_small_linear_bwd_dx(const float* dy0, const float* dy1, const float* dy2, const float* w0, const float* w1, const float* w2,
                             int64_t R, int K, int N, int nout, float* dx, void* stream);
int64_t xdfm_small_linear_bwd_dw_workspace_bytes(int64_t R, int K, int N, int nout);
/* tuning switch: v = rows per thread (1, 2, 4; default 2) of the variant that moves 32 v rows per warp through a shared-memory
 * patch with lane-contiguous 128-bit accesses (single-input launches with K (fwd) / N (dX) in {8, 16, 32} and an output width
 * that is a multiple of 4); 0 = every lane walks its own row */
void xdfm_small_linear_set_staged(int v);
int xdfm_small_linear_bwd_dw(const float* x, const float* dy0, const float* dy1, const float* dy2, int64_t R, int K, int N, int nout,
                             float* dw0, float* dw1, float* dw2, float* db, void* workspace, void* stream);

/* ---- Multi-value (VarLenSparseFeat) pooling: the bag mode of the embedding lookup.
 * Replaces varlen_embedding_lookup + get_varlen_pooling_list (deepctr/inputs.py:141-155, 212-225) and
 * SequencePoolingLayer.forward (deepctr/layers/sequence.py:51-79).  The T positions of a sequence feature are T slots of the fused
 * gather (xdfm_embed_gather); this turns the slot tensor emb [B, S, D] into the field tensor out [B, F, D]: field f owns the slots
 * [slot0[f], slot0[f] + slen[f]) (contiguous, in order, covering [0, S)), mode[f] = XDFM_BAG_SINGLE (copy; slen 1) / SUM / MEAN / MAX.
 * Mask of position j: lencol[f] < 0 -> ids[b, slot0[f] + j] != 0 (supports_masking=True; 'mean' divides by the number of valid
 * positions + 1e-8); lencol[f] >= 0 -> j < lens[b, lencol[f]] ('mean' divides by that length + 1e-8, as given).  MAX reduces
 * x - (1 - mask) * 1e9 and records the winning position per element in argmax [B, F, D] (required when any field is MAX; first
 * position wins ties).  ids int32 [B, S] are the ids the slot tensor was gathered with; lens int32 [B, nlen] or NULL.
 * den float [B, F] (required when any field is MEAN): the forward stores each MEAN field's divisor there for the backward.
 * slot0 / slen / mode / lencol are HOST arrays [F]; S, F <= 64.  bwd: demb [B, S, D] = d(sum out * dout)/d emb. */
#define XDFM_BAG_SINGLE 0
#define XDFM_BAG_SUM 1
#define XDFM_BAG_MEAN 2
#define XDFM_BAG_MAX 3
int xdfm_bag_pool_fwd(const float* emb, const int32_t* ids, const int32_t* lens, int nlen, int64_t B, int S, int D, int F,
                      const int32_t* slot0, const int32_t* slen, const int32_t* mode, const int32_t* lencol, float* out,
                      int32_t* argmax, float* den, void* stream);
int xdfm_bag_pool_bwd(const float* dout, const int32_t* ids, const int32_t* lens, int nlen, const int32_t* argmax, const float* den,
                      int64_t B, int S, int D, int F, const int32_t* slot0, const int32_t* slen, const int32_t* mode,
                      const int32_t* lencol, float* demb, void* stream);

/* ---- xDeepFM Pro: Supervised-Feature-Generation loss (deepctr/xdeepfm_pro/sfg_decoder.py:266-309).
 * row_w[b] = mask_b / num (mask = label == 1, num = sum(mask) + 1e-8 when positive_only; else 1 / B);
 * masked_ce:  row_loss[r] = row_w[r] * CE(logits[r, :V], targets[r * target_stride]), dlogits = d(sum row_loss)/d logits;
 * masked_mse: row_loss[r] = row_w[r] * mean_j (pred - target)^2, dpred likewise. */
int xdfm_sfg_row_weights(const float* labels, int64_t B, int positive_only, float* row_w, void* stream);
int xdfm_masked_ce(const float* logits, const int32_t* targets, int64_t target_stride, const float* row_w, int64_t R, int V,
                   float* row_loss, float* dlogits, void* stream);
int xdfm_masked_mse(const float* pred, const float* target, const float* row_w, int64_t R, int nd, float* row_loss, float* dpred,
                    void* stream);

/* ---- xDeepFM Pro: AutoDis soft-bucket encoder of the dense features (deepctr/xdeepfm_pro/autodis.py:63-69, 100-127):
 * per dense feature f and sample b, with v = x[b, f]:
 *   h = LeakyReLU_0.2(w1[f,:] * v + b1[f,:]);  score = W2[f] h + b2[f];  p = softmax(score / temp[f]);  out[b, f, :] = p @ meta[f]
 * x [B, nd]; w1, b1, b2 [nd, nb]; W2 [nd, nb, nb]; meta [nd, nb, E]; temp [nd]; out [B, nd*E].
 * bwd: dout [B, nd*E] -> gpack [nd, xdfm_autodis_param_count(nb, E)], per feature in the order meta | W2 | b2 | w1 | b1 | temp
 * (deterministic two-stage reduction over the batch); workspace sized by xdfm_autodis_bwd_workspace_bytes. */
int xdfm_autodis_fwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, const float* meta,
                     const float* temp, int64_t B, int nd, int nb, int E, float* out, void* stream);
int64_t xdfm_autodis_param_count(int nb, int E);
int64_t xdfm_autodis_bwd_workspace_bytes(int64_t B, int nd, int nb, int E);
int xdfm_autodis_bwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, const float* meta,
                     const float* temp, const float* dout, int64_t B, int nd, int nb, int E, float* gpack, void* workspace,
                     void* stream);

/* ---- single-node multi-GPU: row-sharded tables over NVLink peer memory (no reference equivalent: the reference replicates whole
 * tables under nn.DataParallel, deepctr/models/basemodel.py:206-209, deepctr/inputs.py:167-180).
 * Global row r of a table lives on rank r % G at local row r / G; a rank keeps its shards of all tables of a set in ONE buffer
 * [rows_g, width]; feature f of rank g starts at local row feat_base[g*m + f] of buffer g.
 * xdfm_ipc_*: cudaMalloc'ed (zero-filled) memory that other processes of the node can map (cudaIpc handles are 64 bytes). */
int xdfm_ipc_alloc(int64_t bytes, void** dptr);
int xdfm_ipc_free(void* dptr);
int xdfm_ipc_export(void* dptr, void* handle64);
int xdfm_ipc_open(const void* handle64, void** dptr);
int xdfm_ipc_close(void* dptr);
/* forward: out_emb[b,f,:] = emb_shards[id % G][feat_base[(id % G)*m + f] + id / G, :], out_lin as xdfm_embed_gather.
 * emb_shards_dev / lin_shards_dev: DEVICE arrays [G] of (peer-mapped) buffer pointers; feat_base_dev: DEVICE int64 [G*m];
 * vocab: HOST [m].  Either output may be NULL. */
int xdfm_embed_gather_sharded(const float* const* emb_shards_dev, const float* const* lin_shards_dev, const int64_t* feat_base_dev,
                              const int32_t* vocab, const int32_t* ids, int64_t B, int m, int D, int G, float* out_emb,
                              const float* dense, int nd, const float* dense_w, float* out_lin, void* stream);
/* backward, batch side: keys = owner*key_stride + local row, sorted owner-major -> run-length segments (outputs as
 * xdfm_embed_bwd_segments) + owner_ranges [G+1]: rank g's segments are [owner_ranges[g], owner_ranges[g+1]). */
int64_t xdfm_shard_workspace_bytes(int64_t n_keys);
int xdfm_shard_segments(const int32_t* ids, int64_t B, int m, int G, uint32_t key_stride, const int64_t* feat_base_dev,
                        const int32_t* vocab, void* workspace, int64_t workspace_bytes, uint32_t* uniq_keys, int32_t* seg_offsets,
                        int32_t* sorted_pos, int32_t* num_segments, int32_t* owner_ranges, void* stream);
/* backward, owner side: pull this rank's ranges (keys, row sums [.,D], first-order sums [.]) from every peer's exchange buffers
 * (HOST arrays [G] of peer-mapped device pointers), concatenate them in rank order into rows / rows_lin ([n_cap, D] / [n_cap]),
 * stable-sort + run-length encode the local keys: uniq_keys (local rows), seg_offsets [n_cap+1], sorted_pos [n_cap],
 * num_segments.  n_cap = capacity (sum over peers of their batch keys); workspace of xdfm_shard_workspace_bytes(n_cap). */
int xdfm_shard_pull_segments(const void* const* peer_keys, const void* const* peer_gsum, const void* const* peer_gsum_lin,
                             const void* const* peer_ranges, int G, int rank, uint32_t key_stride, int D, int64_t n_cap,
                             void* workspace, int64_t workspace_bytes, float* rows, float* rows_lin, uint32_t* uniq_keys,
                             int32_t* seg_offsets, int32_t* sorted_pos, int32_t* num_segments, void* stream);

/* ---- tcgen05 self-test (diagnostic): D[128,N] = A[128,K] * B[N,K]^T, bf16 in / fp32 out, one CTA.
 * mode 0: A via TMA + shared-memory descriptor (SS); mode 1: A stored to TMEM by the threads (TS, the CIN operand path). */
int xdfm_tc_selftest_gemm(const void* A, const void* Bm, int N, int K, int mode, float* out, void* stream);
/* profiling only: SM-cycle latencies of the hand-off primitives (idle tcgen05.commit, mbarrier arrive, 13 MMAs + commit, their issue
 * time, try_wait on a completed phase, tcgen05.ld + wait, round trips by arrive / by commit against one lane and against eight
 * warps) into out[16] (device int64) */
int xdfm_tc_latency_probe(long long* out, int setmaxnreg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XDFM_H_ */
