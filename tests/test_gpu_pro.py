"""xDeepFM Pro (SFG auxiliary decoder) on the GPU against fixtures produced by the unmodified reference
(oracle/make_golden.py pro) and against torch for the masked-loss kernels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepctr import ops
from tests.helpers import PRO_CASES, assert_close, build_product_model, golden_grads, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("R,V,positive_only", [(37, 50, True), (5, 3, False), (64, 5000, True), (1, 1, True), (16, 257, False)])
def test_masked_cross_entropy_kernel(R, V, positive_only):
    g = torch.Generator().manual_seed(R + V)
    logits = 3 * torch.randn(R, V, generator=g)
    ids = torch.randint(0, V, (R, 3), generator=g).to(torch.int32)
    labels = (torch.rand(R, generator=g) < 0.4).float()
    mask = (labels == 1).double() if positive_only else torch.ones(R, dtype=torch.float64)
    num = mask.sum() + 1e-8 if positive_only else R
    ld = logits.double().requires_grad_(True)
    ref = (F.cross_entropy(ld, ids[:, 1].long(), reduction="none") * mask).sum() / num
    ref.backward()
    lg = logits.to(DEV).requires_grad_(True)
    row_w = ops.sfg_row_weights(labels.to(DEV), positive_only)
    assert_close(row_w, mask / num, 1e-6, 1e-9, "row weights")
    out = ops.MaskedCE.apply(lg, ids.to(DEV), 1, row_w)
    (2.5 * out).backward()
    assert_close(out.reshape(()), ref.detach(), 2e-5, 1e-7, "masked CE")
    assert_close(lg.grad, 2.5 * ld.grad, 2e-5, 2e-5 * float(ld.grad.abs().max()) + 1e-12, "d logits")


def test_masked_mse_kernel():
    g = torch.Generator().manual_seed(0)
    pred, tgt = torch.randn(33, 3, generator=g), torch.rand(33, 3, generator=g)
    labels = (torch.rand(33, generator=g) < 0.4).float()
    mask = (labels == 1).double()
    pd = pred.double().requires_grad_(True)
    ref = (((pd - tgt.double()) ** 2).mean(-1) * mask).sum() / (mask.sum() + 1e-8)
    ref.backward()
    pg = pred.to(DEV).requires_grad_(True)
    out = ops.MaskedMSE.apply(pg, tgt.to(DEV), ops.sfg_row_weights(labels.to(DEV), True))
    out.backward()
    assert_close(out.reshape(()), ref.detach(), 2e-5, 1e-7, "masked MSE")
    assert_close(pg.grad, pd.grad, 2e-5, 1e-8, "d pred")


@pytest.mark.parametrize("name", PRO_CASES)
def test_pro_train_step_matches_reference_fixture(name):
    spec, params, z = load_case(name)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    model.train()
    y_pred, info = model.forward_with_sfg(X, y)
    y_pred = y_pred.squeeze()
    assert_close(y_pred, z["y_pred"], 2e-5, 2e-6, "y_pred")
    if spec.use_sfg:
        sfg = info["sfg_loss"]
        assert abs(sfg.item() - float(z["sfg_loss"])) <= 2e-5 * abs(float(z["sfg_loss"]))
    else:
        assert info is None
        sfg = torch.zeros((), device=DEV)
    loss = F.binary_cross_entropy(y_pred, y, reduction="sum")
    total = loss + model.get_regularization_loss() + model.aux_loss + model.sfg_weight * sfg
    assert abs(total.item() - float(z["total"].item())) <= 2e-5 * abs(float(z["total"].item()))
    model.zero_grad()
    total.backward()
    named = dict(model.named_parameters())
    for k, g in golden_grads(z).items():
        got = named[k].grad
        got = torch.zeros_like(named[k]) if got is None else got
        assert_close(got, g, 2e-4, 2e-5 * max(g.abs().max().item(), 1e-6), "grad " + k)
    model.eval()
    with torch.no_grad():
        y_eval, info = model.forward_with_sfg(X, y)
        assert info is None                       # no SFG term in eval mode (SURVEY.md 3.5 [probed])
        assert_close(y_eval, z["y_pred_eval"], 2e-5, 2e-6, "eval y_pred")


def test_pro_fused_sgd_step_equals_gradient_step():
    spec, params, z = load_case("pro_small")
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile("sgd", "binary_crossentropy")
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    ids, dense = model.split_input(X)
    accum = torch.zeros(1, dtype=torch.float64, device=DEV)
    model.train()
    model.train_step(ids, dense, y, accum)
    assert abs(accum.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    assert abs(model._sfg_accum.item() - float(z["sfg_loss"])) <= 2e-5 * abs(float(z["sfg_loss"]))
    grads = golden_grads(z)
    for k, p in model.state_dict().items():
        expect = params[k].double() - 0.01 * grads[k].double()
        scale = (0.01 * grads[k]).abs().max().item()
        assert_close(p, expect, 0, 3e-4 * scale + 1e-7 * params[k].abs().max().item(), "sgd step " + k)


def test_pro_fit_trajectory_matches_reference():
    spec, params, z = load_case("fit_pro_small_adam")
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile("adam", "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    for g in model.optim.param_groups:
        g["lr"] = float(z["lr"])
    X, y = z["X"], z["y"]
    names = list(model.feature_index.keys())
    xd = {n: X[:, i].copy() for i, n in enumerate(names)}
    hist = model.fit(xd, y.reshape(-1, 1), batch_size=int(z["batch_size"]), epochs=int(z["epochs"]), verbose=0, shuffle=False,
                     validation_data=(dict(xd), y.reshape(-1, 1)))
    assert np.allclose(hist.history["loss"], z["history_loss"], rtol=1e-3), (hist.history["loss"], z["history_loss"])
    assert np.allclose(hist.history["sfg_loss"], z["history_sfg_loss"], rtol=1e-3), (hist.history["sfg_loss"], z["history_sfg_loss"])
    assert np.allclose(hist.history["val_auc"], z["history_val_auc"], atol=5e-3)
    pred = model.predict(dict(xd), batch_size=int(z["batch_size"]))
    assert np.allclose(pred, z["pred"], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("B,nd,nb,E", [(1, 1, 16, 8), (129, 3, 16, 8), (300, 13, 16, 32), (40, 2, 5, 4), (257, 4, 32, 64), (64, 5, 8, 16)])
def test_autodis_kernel_matches_torch(B, nd, nb, E):
    """ops.AutoDis (one fused launch forward, two backward) against the reference's op sequence in fp64 torch
    (autodis.py:100-127): outputs 1e-5, every parameter gradient 2e-5 of its scale; gradients bit-identical run to run."""
    from deepctr.xdeepfm_pro.autodis import AutoDisLayer
    torch.manual_seed(B + nb)
    layer = AutoDisLayer(nd, num_buckets=nb, embedding_dim=E, temperature=0.7, device=DEV)
    with torch.no_grad():
        layer.meta_embeddings.mul_(30.0)
        layer.feature_temperatures.copy_(0.5 + torch.rand(nd, device=DEV))
    x = torch.rand(B, nd, device=DEV) * 2 - 0.5
    gout = torch.randn(B, nd * E, device=DEV)
    flat, per = layer(x)
    assert flat.shape == (B, nd * E) and len(per) == nd and per[0].shape == (B, 1, E)
    (flat * gout).sum().backward()
    got = {k: p.grad.detach().clone() for k, p in layer.named_parameters()}
    # reference op sequence, float64
    p64 = {k: p.detach().double().cpu().requires_grad_(True) for k, p in layer.named_parameters()}
    xd = x.double().cpu()
    outs = []
    for f in range(nd):
        h = F.leaky_relu(xd[:, f:f + 1] @ p64["bucket_projectors.%d.0.weight" % f].t() + p64["bucket_projectors.%d.0.bias" % f], 0.2)
        sc = h @ p64["bucket_projectors.%d.2.weight" % f].t() + p64["bucket_projectors.%d.2.bias" % f]
        outs.append(torch.softmax(sc / p64["feature_temperatures"][f], dim=-1) @ p64["meta_embeddings"][f])
    ref = torch.cat(outs, dim=-1)
    (ref * gout.double().cpu()).sum().backward()
    assert_close(flat, ref.detach(), 1e-5, 1e-6, "autodis out")
    for k, g in got.items():
        r = p64[k].grad
        assert_close(g, r, 2e-5, 2e-5 * max(r.abs().max().item(), 1e-9), "autodis grad " + k)
    layer.zero_grad()
    flat2, _ = layer(x)
    (flat2 * gout).sum().backward()
    assert torch.equal(flat2, flat)
    for k, p in layer.named_parameters():
        assert torch.equal(p.grad, got[k]), "autodis gradient not reproducible: " + k


def test_autodis_rejects_unsupported_shapes_and_cpu():
    from deepctr.xdeepfm_pro.autodis import AutoDisLayer
    layer = AutoDisLayer(2, num_buckets=64, embedding_dim=256, device=DEV)
    with pytest.raises(RuntimeError):
        layer(torch.rand(4, 2, device=DEV))                  # does not fit the fused kernel: explicit error, no fallback
    cpu = AutoDisLayer(2, num_buckets=4, embedding_dim=4, device="cpu")
    with pytest.raises(RuntimeError):
        cpu(torch.rand(4, 2))
    with pytest.raises(ValueError):
        AutoDisLayer(2, num_buckets=4, embedding_dim=4, device=DEV)(torch.rand(4, 3, device=DEV))


def test_pro_autodis_fit_trajectory_matches_reference():
    """fit() of xDeepFMPro(use_autodis=True) (CUDA-graph replayed SFG step) against the reference's History."""
    spec, params, z = load_case("fit_pro_autodis_adam")
    model = build_product_model(spec, DEV)
    assert set(model.state_dict().keys()) == set(params.keys())
    model.load_state_dict(params, strict=True)
    model.compile("adam", "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    for g in model.optim.param_groups:
        g["lr"] = float(z["lr"])
    X, y = z["X"], z["y"]
    names = list(model.feature_index.keys())
    xd = {n: X[:, i].copy() for i, n in enumerate(names)}
    hist = model.fit(xd, y.reshape(-1, 1), batch_size=int(z["batch_size"]), epochs=int(z["epochs"]), verbose=0, shuffle=False,
                     validation_data=(dict(xd), y.reshape(-1, 1)))
    assert np.allclose(hist.history["loss"], z["history_loss"], rtol=1e-3), (hist.history["loss"], z["history_loss"])
    assert np.allclose(hist.history["sfg_loss"], z["history_sfg_loss"], rtol=1e-3), (hist.history["sfg_loss"], z["history_sfg_loss"])
    assert np.allclose(hist.history["val_auc"], z["history_val_auc"], atol=5e-3)
    pred = model.predict(dict(xd), batch_size=int(z["batch_size"]))
    assert np.allclose(pred, z["pred"], rtol=2e-3, atol=2e-4)
    final = {k[len("final::"):]: z[k] for k in z.files if k.startswith("final::")}
    sd = model.state_dict()
    for k in ("autodis_encoder.autodis.meta_embeddings", "autodis_encoder.autodis.feature_temperatures",
              "autodis_encoder.autodis.bucket_projectors.1.2.weight"):
        assert_close(sd[k], final[k], 2e-3, 2e-3 * float(np.abs(final[k]).max()), "trained " + k)


def test_pro_cuda_graph_replay_equals_eager_steps():
    """The SFG step (decoder, per-field heads, masked losses) replayed from a CUDA graph leaves the same weights and losses."""
    spec, params, z = load_case("pro_small")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    batches = [(X[i * 10:(i + 1) * 10], y[i * 10:(i + 1) * 10]) for i in range(4)] * 2
    out = []
    for graph in (False, True):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("adam", "binary_crossentropy")
        model.use_cuda_graph = graph
        model.train()
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        for Xb, yb in batches:
            ids, dense = model.split_input(Xb.to(DEV))
            model.train_step(ids, dense, yb.to(DEV), accum)
        assert bool(model._graphs) == graph
        out.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, accum.item(), model._sfg_accum.item()))
    (sd_e, loss_e, sfg_e), (sd_g, loss_g, sfg_g) = out
    for k in sd_e:
        assert torch.equal(sd_e[k], sd_g[k]), k
    assert abs(loss_e - loss_g) <= 1e-9 * abs(loss_e) and abs(sfg_e - sfg_g) <= 1e-9 * abs(sfg_e)


@pytest.mark.parametrize("graph", [False, True])
def test_pro_positive_rows_only_sfg_pass_equals_all_rows(graph):
    """train_on_batch / fit know the batch's labels on the host and run the SFG decoder, heads and masked losses on the label-1
    rows only (bucketed static shapes, device-side stable sort, no sync); rows with weight 0 contribute nothing, so weights and
    losses must equal the all-rows step up to summation order."""
    spec, params, z = load_case("pro_small")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    out = []
    for hinted in (False, True):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("sgd", "binary_crossentropy")          # linear in the gradient: summation-order noise stays noise
        model.use_cuda_graph = graph
        model.SFG_ROW_BUCKET = 4
        model.train()
        losses = []
        for rep in range(4):                       # same shapes four times: the graph path captures on the third call
            for lo in (0, 20):
                Xb, yb = X[lo:lo + 20], y[lo:lo + 20]
                ids, dense = model.split_input(Xb.to(DEV))
                if hinted:
                    losses.append(model.train_on_batch(ids.cpu(), dense.cpu(), yb))
                    assert model._npos_hint is None
                else:
                    accum = torch.zeros(1, dtype=torch.float64, device=DEV)
                    model.train_step(ids, dense, yb.to(DEV), accum)
                    losses.append(accum.item())
        if hinted:
            npos = int((y[:20] == 1).sum())
            model._host_label_hint(y[:20])
            assert model._sfg_rows_cap(20) == min(20, max(4, -(-npos // 4) * 4))
            model._npos_hint = None
            if graph:
                assert model._graphs
        out.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, losses, model._sfg_accum.item()))
    (sd_a, loss_a, sfg_a), (sd_b, loss_b, sfg_b) = out
    assert np.allclose(loss_a, loss_b, rtol=1e-5)
    assert abs(sfg_a - sfg_b) <= 1e-5 * abs(sfg_a)
    for k in sd_a:
        assert_close(sd_b[k], sd_a[k], 2e-4, 2e-5 * max(sd_a[k].abs().max().item(), 1e-6), "positive-rows-only step: " + k)
