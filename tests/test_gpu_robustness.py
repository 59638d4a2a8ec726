"""Host-side hazards around the fused step (each one was a review finding): recycled device addresses in the segment cache,
the reference's eval-mode-after-validation quirk, pickling a model that holds CUDA graphs, gradients that exist before the flat
views, workspaces that grow after a graph was captured; plus optimizer-state checkpoints."""
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepctr import ops
from tests.helpers import assert_close, build_product_model, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _table_grads(model, X, y):
    """One function-scoped forward / backward: every local (the split ids included) is freed on return."""
    y_pred = model(X).squeeze()
    loss = F.binary_cross_entropy(y_pred, y, reduction="sum")
    model.zero_grad()
    loss.backward()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if "embedding_dict" in k}


def test_segment_cache_survives_recycled_ids_addresses():
    """model(X) + backward in a user loop: split_input() returns a fresh ids tensor per call, which the allocator places at the
    address the previous call's ids had.  The second step must scatter its gradients to ITS rows."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    Xa, ya, Xb, yb = X[:32].contiguous(), y[:32].contiguous(), X[32:].contiguous(), y[32:].contiguous()
    assert not torch.equal(Xa, Xb)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.train()
    _table_grads(model, Xa, ya)
    second = _table_grads(model, Xb, yb)          # same shapes, previous ids freed: same device address
    fresh = build_product_model(spec, DEV)
    fresh.load_state_dict(params, strict=True)
    fresh.train()
    expect = _table_grads(fresh, Xb, yb)
    for k in expect:
        assert torch.equal(second[k], expect[k]), k
    # the stand-alone first-order module keeps its own cache
    lin = model.linear_model
    for Xi in (Xa, Xb):
        model.zero_grad()
        lin(Xi).sum().backward()
    got = {k: p.grad.clone() for k, p in lin.named_parameters() if "embedding_dict" in k}
    fresh.zero_grad()
    fresh.linear_model(Xb).sum().backward()
    for k, p in fresh.linear_model.named_parameters():
        if "embedding_dict" in k:
            assert torch.equal(got[k], p.grad), k


def test_fit_with_validation_stays_on_the_graph_and_lazy_path_after_the_first_epoch():
    """The reference's fit() leaves the model in eval mode after the first validation pass (kept for result parity).  That must not
    push the following epochs onto eager launches or onto a full table flush per step."""
    spec, params, z = load_case("fit_small_adam")
    X, y = z["X"], z["y"]
    losses = []
    for graph in (True, False):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("adam", "binary_crossentropy", metrics=["auc"])
        model.use_cuda_graph = graph
        names = list(model.feature_index.keys())
        xd = {n: X[:, i].copy() for i, n in enumerate(names)}
        flushes, eager = [0], [0]
        orig_flush, orig_eager = model.optim.flush, model._train_step_eager

        def counting_flush(_orig=orig_flush):
            flushes[0] += 1
            return _orig()

        def counting_eager(*a, _orig=orig_eager, _m=model, **k):
            if not _m._capturing:
                eager[0] += 1
            return _orig(*a, **k)

        model.optim.flush = counting_flush
        model._train_step_eager = counting_eager
        hist = model.fit(xd, y.reshape(-1, 1), batch_size=8, epochs=4, verbose=0, shuffle=False,
                         validation_data=(dict(xd), y.reshape(-1, 1)))
        steps = 4 * 12
        assert not model.training                      # the quirk itself is kept
        if graph:
            # eager steps: two warm-ups per (shape, mode) key -- train mode in epoch 1, eval mode from epoch 2 on
            assert eager[0] <= 6, "eager steps: %d of %d" % (eager[0], steps)
        # flush(): epoch end (pop_reg_loss), predict / evaluate, state checks -- a handful per epoch, never one per step
        assert flushes[0] <= 4 * 5, "table flushes: %d for %d steps" % (flushes[0], steps)
        losses.append(hist.history["loss"])
    assert np.allclose(losses[0], losses[1], rtol=1e-9), losses


def test_torch_save_of_the_whole_model_after_a_graphed_fit_and_resume():
    """ModelCheckpoint's default (save_weights_only=False, as in the reference: callbacks.py:58-61) pickles the model.  After graph
    capture the model holds CUDAGraph objects; they are dropped from the pickle, and the reloaded model (optimizer moments
    included) continues exactly like the original."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile("adam", "binary_crossentropy")
    model.train()
    ids, dense = model.split_input(X)
    accum = torch.zeros(1, dtype=torch.float64, device=DEV)
    for _ in range(4):
        model.train_step(ids, dense, y, accum)
    assert model._graphs
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    clone = torch.load(buf, weights_only=False)
    assert not clone._graphs and clone.optim.steps == 4
    for m in (model, clone):
        m.train()
        for _ in range(3):
            m.train_step(ids, dense, y, accum)
    sd_a, sd_b = model.state_dict(), clone.state_dict()
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k


@pytest.mark.parametrize("optimizer", ["adam", "adagrad", "rmsprop", "sgd"])
def test_optimizer_state_dict_resumes_bit_identically(optimizer):
    """train 6 steps == train 3, save model + optimizer state_dict, load both into a NEW model, train 3 more."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    batches = [(X[i * 16:(i + 1) * 16].contiguous(), y[i * 16:(i + 1) * 16].contiguous()) for i in range(4)]

    def make():
        m = build_product_model(spec, DEV)
        m.load_state_dict(params, strict=True)
        m.compile(optimizer, "binary_crossentropy")
        m.train()
        return m

    def run(m, lo, hi):
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        for s in range(lo, hi):
            Xb, yb = batches[s % 4]
            ids, dense = m.split_input(Xb)
            m.train_step(ids, dense, yb, accum)
        return accum.item()

    ref = make()
    run(ref, 0, 6)
    a = make()
    run(a, 0, 3)
    buf = io.BytesIO()
    torch.save({"model": a.state_dict(), "optim": a.optim.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, map_location="cpu")
    b = make()
    b.load_state_dict(ck["model"])
    b.optim.load_state_dict(ck["optim"])
    assert b.optim.steps == 3
    run(b, 3, 6)
    sd_ref, sd_b = ref.state_dict(), b.state_dict()
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_b[k]), k
    st_ref, st_b = ref.optim.state_dict()["fused"], b.optim.state_dict()["fused"]
    for name, (s1, s2) in st_ref["dense"].items():
        for u, v in zip((s1, s2), st_b["dense"][name]):
            assert (u is None and v is None) or torch.equal(u, v), name
    for ta, tb in zip(st_ref["tables"], st_b["tables"]):
        for key in ("s1", "s2"):
            if ta[key] is not None:
                for u, v in zip(ta[key], tb[key]):
                    assert torch.equal(u, v)


def test_first_step_of_a_user_loop_keeps_the_gradients_computed_before_the_flat_views_exist():
    """Generic path with a named (fused) optimizer: forward, backward, THEN model.optim.step() for the first time.  prepare()
    rebinds every dense .grad to a view of the flat gradient buffer; the values backward left must move with it."""
    spec, params, z = load_case("xdeepfm_small")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile("sgd", "binary_crossentropy")
    model.train()
    y_pred = model(X).squeeze()
    total = F.binary_cross_entropy(y_pred, y, reduction="sum") + model.get_regularization_loss()
    model.optim.zero_grad()
    total.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    assert grads["dnn.linears.0.weight"].abs().max().item() > 0
    model.optim.step()
    for k, p in model.named_parameters():
        assert_close(p, before[k] - 0.01 * grads[k], 0, 1e-6 * before[k].abs().max().item() + 1e-9, "sgd step " + k)


def test_graphs_captured_before_a_workspace_grew_are_not_replayed():
    """fit(bs=16) -> bs=64 -> bs=16 in one process: the larger batch replaces cached workspaces whose addresses the first graph
    holds.  Results must equal an all-eager run."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    plan = [16] * 4 + [64] * 4 + [16] * 4
    out = []
    for graph in (False, True):
        ops._WS.clear()
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("adam", "binary_crossentropy")
        model.use_cuda_graph = graph
        model.train()
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        for s, bs in enumerate(plan):
            lo = (s * 16) % (64 - bs + 1)
            ids, dense = model.split_input(X[lo:lo + bs].contiguous())
            model.train_step(ids, dense, y[lo:lo + bs].contiguous(), accum)
        out.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, accum.item()))
    for k in out[0][0]:
        assert torch.equal(out[0][0][k], out[1][0][k]), k
    assert abs(out[0][1] - out[1][1]) <= 1e-9 * abs(out[0][1])
