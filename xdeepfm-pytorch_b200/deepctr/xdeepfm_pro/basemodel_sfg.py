"""BaseModel + Supervised Feature Generation (reference: deepctr/xdeepfm_pro/basemodel_sfg.py:96-677).

The reference copies the whole BaseModel and adds `forward_with_sfg`, `compute_sfg_loss` and the `+ sfg_weight * sfg_loss` term
(basemodel_sfg.py:316-349); here the B200 BaseModel is extended instead: the fused train step back-propagates the BCE gradient
and the SFG loss in one autograd pass, the SFG loss is accumulated on the device and reported as `History['sfg_loss']`."""
import torch

from .. import ops
from ..inputs import DenseFeat, SparseFeat
from ..models.basemodel import BaseModel
from .sfg_decoder import SFGDecoder, SFGLoss


class BaseModelSFG(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, l2_reg_linear=1e-5, l2_reg_embedding=1e-5, init_std=0.0001,
                 seed=1024, task='binary', device='cpu', gpus=None, use_sfg=True, sfg_weight=0.1, sfg_hidden_units=(128, 64),
                 sfg_dropout=0.1, sfg_positive_only=True, sfg_use_label_attention=True):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=l2_reg_linear,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task, device=device, gpus=gpus)
        if self._all_varlen and use_sfg:
            # the reference's SFG decoder is sized from the SparseFeat columns only (sfg_decoder.py:52-77) and its forward_with_sfg
            # fails with a shape error once a pooled sequence embedding joins the list (probed); without SFG the model is fine
            raise NotImplementedError("xDeepFMPro(use_sfg=True) with VarLenSparseFeat columns: the reference's SFG decoder does not "
                                      "accept multi-value columns either (its input width ignores them); pass use_sfg=False")
        self.use_sfg, self.sfg_weight, self.sfg_positive_only = use_sfg, sfg_weight, sfg_positive_only
        self.sparse_feature_columns = [fc for fc in dnn_feature_columns if isinstance(fc, SparseFeat)] if dnn_feature_columns else []
        self.dense_feature_columns = [fc for fc in dnn_feature_columns if isinstance(fc, DenseFeat)] if dnn_feature_columns else []
        self.embedding_dim = self.sparse_feature_columns[0].embedding_dim if self.sparse_feature_columns else 8
        if use_sfg:
            dims = {fc.name: fc.vocabulary_size for fc in self.sparse_feature_columns}
            dense_names = [fc.name for fc in self.dense_feature_columns]
            self.sfg_decoder = SFGDecoder(self.embedding_dim, dims, dense_names, hidden_units=sfg_hidden_units,
                                          dropout_rate=sfg_dropout, use_label_aware_attention=sfg_use_label_attention, device=device)
            self.sfg_loss_fn = SFGLoss([fc.name for fc in self.sparse_feature_columns], dense_names,
                                       positive_only=sfg_positive_only, device=device)
        else:
            self.sfg_decoder = None
            self.sfg_loss_fn = None
        self._sfg_accum = None
        import os as _os
        if _os.environ.get("XDFM_CUDA_GRAPH_SFG", "1") == "0":
            self.use_cuda_graph = False
        self.to(device)

    # ---- positive rows only -------------------------------------------------------------------------------
    # With sfg_positive_only (the reference default) rows with label != 1 have weight 0 in every SFG term (sfg_decoder.py:266-273,
    # 290, 303): they add nothing to the loss or to any gradient, so the decoder, its per-field heads (2 * 64 * sum(vocab) FLOP and
    # 4 * sum(vocab) bytes of logits PER ROW) and the masked losses only need the positive rows (hard part 7).
    # Their number is data dependent; it is taken from the HOST labels of the batch (fit() / train_on_batch have them before the
    # copy), rounded up to a bucket of SFG_ROW_BUCKET rows so that shapes stay static (a handful of CUDA graphs); the rows are
    # brought to the front with a stable device-side sort -- no device sync.  Rows beyond the true count inside the bucket are
    # negatives with weight 0: exact.  No hint (train_step called directly) = all rows.
    SFG_ROW_BUCKET = 512

    def _host_label_hint(self, y_host):
        self._npos_hint = None
        if self.use_sfg and self.sfg_positive_only and self.training:
            y = torch.as_tensor(y_host).reshape(-1)
            self._npos_hint = (int((y == 1).sum()), int(y.shape[0]))

    def _sfg_rows_cap(self, B):
        """Static number of rows the SFG pass of the coming step runs on (B = all)."""
        hint = getattr(self, "_npos_hint", None)
        if hint is None or hint[1] != B or not self.sfg_positive_only:
            return B
        bucket = self.SFG_ROW_BUCKET
        return min(B, max(bucket, -(-hint[0] // bucket) * bucket))

    def _graph_key(self, ids, dense, y):
        return super()._graph_key(ids, dense, y) + (self._sfg_rows_cap(ids.shape[0]) if self.use_sfg else None,)

    def train_step(self, ids, dense, y, *args, **kwargs):
        if self._sfg_in_step_collective():
            # data parallel: the global positive count is all-reduced HERE, before the (possibly CUDA-graph replayed) step body,
            # into a static device scalar the body multiplies its row weights with -- no collective inside the captured region
            self._set_global_sfg_scale(y)
        try:
            return super().train_step(ids, dense, y, *args, **kwargs)
        finally:
            self._npos_hint = None           # a hint describes exactly one batch (eager, captured or replayed)
            self._sfg_scale_ready = False

    # ---- SFG loss on the split (ids, dense) feed ------------------------------------------------------
    def sfg_loss_ids(self, ids_all, dense_all, emb, labels):
        """SFG loss [scalar tensor] from this step's embeddings `emb` [B, m, D] (sfg_decoder.py:116-157, 266-309); the targets
        are the ids / dense values themselves (basemodel_sfg.py:446-466)."""
        ids = self._select(ids_all, self._dnn_sparse_sel)
        dd = self.dnn_dense(dense_all)
        dec = self.sfg_decoder
        flat = emb.reshape(emb.shape[0], -1)
        dec_in = torch.cat([flat, dd], dim=-1) if dd.shape[1] > 0 else flat
        row_w = ops.sfg_row_weights(labels, self.sfg_positive_only)
        if self._dist is not None and self._dist.world > 1:
            row_w = row_w * self._global_row_weight_scale(labels)
        B = flat.shape[0]
        cap = self._sfg_rows_cap(B)
        if cap < B:
            sel = torch.argsort(labels.reshape(-1), descending=True, stable=True)[:cap]      # label-1 rows first, batch order kept
            dec_in, labels, row_w = dec_in.index_select(0, sel), labels.reshape(-1).index_select(0, sel), row_w.index_select(0, sel)
            ids, dd = ids.index_select(0, sel).contiguous(), dd.index_select(0, sel)
        h = dec.hidden(dec_in, labels)
        fn = self.sfg_loss_fn
        total = None
        for f, fc in enumerate(self.sparse_feature_columns):
            head = dec.sparse_heads[fc.name]
            l = ops.MaskedCE.apply(ops.linear_act(h, head.weight, head.bias, precision=dec.precision), ids, f, row_w)
            total = l if total is None else total + l
        total = torch.zeros(1, device=emb.device) if total is None else fn.sparse_weight * total
        if dec.dense_head is not None and dd.shape[1] > 0:
            pred = ops.linear_act(h, dec.dense_head.weight, dec.dense_head.bias, precision=dec.precision)
            total = total + fn.dense_weight * ops.MaskedMSE.apply(pred, dd, row_w)
        return total.reshape(())

    # ---- data parallel: the SFG normaliser is a property of the GLOBAL batch ---------------------------------
    # The reference divides the masked reconstruction losses by the number of label-1 rows of the batch it was handed -- with
    # `gpus=[...]` that is the whole G * batch_size batch on one device (basemodel_sfg.py:279-282 scales batch_size, :316 calls
    # self.forward_with_sfg, not the DataParallel wrapper).  One process per GPU sees only its slice, so the local weights
    # mask / (n_local + 1e-8) are rescaled to mask / (n_global + 1e-8) with ONE all-reduce of the count (8 bytes); no host sync.
    def _sfg_in_step_collective(self):
        return self.use_sfg and self.sfg_decoder is not None and self.training and self._dist is not None and self._dist.world > 1

    def _set_global_sfg_scale(self, labels):
        dev = torch.device(self.device)
        if getattr(self, "_sfg_scale_buf", None) is None or self._sfg_scale_buf.device != dev:
            self._sfg_scale_buf = torch.ones(1, dtype=torch.float32, device=dev)
        self._sfg_scale_ready = False
        self._sfg_scale_buf.copy_(self._global_row_weight_scale(labels.to(dev)).reshape(1))
        self._sfg_scale_ready = True

    def _global_row_weight_scale(self, labels):
        if getattr(self, "_sfg_scale_ready", False):
            return self._sfg_scale_buf           # set by train_step() for this batch
        y = labels.reshape(-1)
        local = ((y == 1).sum() if self.sfg_positive_only else torch.full((), y.shape[0], device=y.device)).to(torch.float64)
        glob = local.clone().reshape(1)
        self._dist.all_reduce_sum(glob)
        eps = 1e-8 if self.sfg_positive_only else 0.0
        return ((local + eps) / (glob[0] + eps)).to(torch.float32)

    def _empty_step_collectives(self):
        if self._sfg_in_step_collective() and not getattr(self, "_sfg_scale_ready", False):
            zero = torch.zeros(1, dtype=torch.float64, device=torch.device(self.device))
            self._dist.all_reduce_sum(zero)          # this rank has no rows in the last partial batch: 0 positives of 0 rows

    def compute_sfg_loss(self, X, sparse_embedding_list, dense_value_list, labels):
        """Reference-shaped entry (basemodel_sfg.py:420-476)."""
        if not self.use_sfg or self.sfg_decoder is None:
            return torch.tensor(0.0, device=self.device), {}
        ids, dense = self.split_input(X)
        emb = torch.cat(list(sparse_embedding_list), dim=1)
        loss = self.sfg_loss_ids(ids, dense, emb, labels)
        return loss, {'sfg_loss': loss, 'sfg_loss_dict': {'sfg_total': loss.detach()}}

    def forward_with_sfg(self, X, y=None):
        """(y_pred, sfg_info) like the reference (xdeepfm_pro.py:203-274): sfg_info only in training mode with labels."""
        X = X.to(self.device) if not X.is_cuda else X
        ids, dense = self.split_input(X)
        y_pred = self.forward_ids(ids, dense)
        info = None
        if self.use_sfg and y is not None and self.training:
            loss = self.sfg_loss_ids(ids, dense, self._last_emb, y.to(self.device))
            info = {'sfg_loss': loss, 'sfg_loss_dict': {'sfg_total': loss.detach()}}
        self._last_emb = None
        return y_pred, info

    def forward(self, X):
        return self.forward_with_sfg(X, None)[0]

    # ---- fused train step: BCE gradient and sfg_weight * SFG loss back-propagated together ---------------
    def _train_step_inner(self, opt, ids, dense, y, loss_accum, pred_log, pred_off):
        y_pred = self.forward_ids(ids, dense)
        yv = y.reshape(-1)
        roots, grads = [], []
        if self._loss_name == "binary_crossentropy":
            _, dy = ops.bce_sum(y_pred, yv, loss_accum)
            roots.append(y_pred)
            grads.append(dy.view_as(y_pred))
        else:
            loss = self.loss_func(y_pred.reshape(-1), yv, reduction="sum")
            loss_accum += loss.detach().double()
            roots.append(loss)
            grads.append(torch.ones_like(loss))
        if self.use_sfg and self.sfg_decoder is not None and self.training:      # gated on training mode: xdeepfm_pro.py:265
            sfg = self.sfg_loss_ids(ids, dense, self._last_emb, yv)
            if self._sfg_accum is None or self._sfg_accum.device != sfg.device:
                self._sfg_accum = torch.zeros(1, dtype=torch.float64, device=sfg.device)
            self._sfg_accum += sfg.detach().double()
            roots.append(sfg)
            grads.append(torch.full_like(sfg, float(self.sfg_weight)))
        torch.autograd.backward(roots, grads)
        # nothing may keep this step's autograd graph alive: a surviving graph pins its AccumulateGrad nodes (bound to the stream
        # they were created on), the next step would reuse them, and under CUDA-graph capture (side stream) autograd would then
        # synchronise with the uncaptured default stream -> cudaErrorStreamCaptureIsolation
        self._last_emb = None
        del roots, grads
        self._optimizer_phases(opt)
        y_pred = y_pred.detach()
        if pred_log is not None:
            pred_log[pred_off:pred_off + yv.shape[0]] = y_pred.reshape(-1)
        return y_pred

    def _epoch_begin(self):
        if self._sfg_accum is not None:
            self._sfg_accum.zero_()

    def _epoch_extra(self, sample_num):
        """(extra total-loss sum, extra History entries): total += sfg_weight * sum(sfg), History['sfg_loss'] = sum(sfg) / N
        (basemodel_sfg.py:344, 369-371)."""
        if not self.use_sfg:
            return 0.0, {}
        acc = self._sfg_accum
        if self._dist is not None and self._dist.world > 1:
            acc = torch.zeros(1, dtype=torch.float64, device=torch.device(self.device)) if acc is None else acc.clone()
            self._dist.all_reduce_sum(acc)           # every rank accumulated its share of the globally normalised loss
        s = 0.0 if acc is None else float(acc.item())
        return self.sfg_weight * s, {"sfg_loss": s / sample_num}
