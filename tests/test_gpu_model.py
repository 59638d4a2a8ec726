"""GPU parity of the whole model against the fixtures produced by the unmodified reference (tests/golden) and the
CPU oracle: predictions, loss, every parameter gradient, optimizer trajectories of fit()."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import xdeepfm_oracle as O
from tests.helpers import FIT_CASES, assert_close, build_product_model, golden_gradnorms, golden_grads, golden_gradsamples, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
XDEEPFM_CASES = ["xdeepfm_small", "xdeepfm_small_nosplit", "xdeepfm_small_linearact", "xdeepfm_small_nodense",
                 "xdeepfm_small_zipf", "xdeepfm_cfg1", "xdeepfm_cfg2", "attn_small", "attn_small_3heads", "attn_v2_small"]


@pytest.mark.parametrize("name", XDEEPFM_CASES)
def test_forward_backward_matches_reference_fixture(name):
    spec, params, z = load_case(name)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    X, y = torch.from_numpy(z["X"]).to(DEV), torch.from_numpy(z["y"]).to(DEV)
    model.train()
    y_pred = model(X).squeeze()
    # fp32 end to end (CUDA-core CIN path): tolerance 2e-5 relative on probabilities
    assert_close(y_pred, z["y_pred"], 2e-5, 2e-6, "y_pred")
    loss = F.binary_cross_entropy(y_pred, y, reduction="sum")
    reg = model.get_regularization_loss()
    total = loss + reg + model.aux_loss
    assert abs(loss.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    assert abs(reg.item() - float(z["reg_loss"])) <= 2e-5 * abs(float(z["reg_loss"])) + 1e-9
    model.zero_grad()
    total.backward()
    named = dict(model.named_parameters())
    for k, g in golden_grads(z).items():
        got = named[k].grad
        assert got is not None, k
        assert_close(got, g, 2e-4, 2e-5 * max(g.abs().max().item(), 1e-6), "grad " + k)
    for k, n in golden_gradnorms(z).items():
        got = named[k].grad
        gn = 0.0 if got is None else got.double().norm().item()
        assert abs(gn - n) <= 2e-4 * max(n, 1e-6), "grad norm " + k
    for k, (stride, g) in golden_gradsamples(z).items():
        assert_close(named[k].grad.flatten()[::stride], g, 2e-4, 2e-5 * max(g.abs().max().item(), 1e-6), "grad sample " + k)
    model.eval()
    with torch.no_grad():
        assert_close(model(X), z["y_pred_eval"], 2e-5, 2e-6, "eval y_pred")


def test_fused_step_gradients_match_oracle():
    """The fused path (sparse stash + flat dense gradient buffer) must see the same gradients as the oracle."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile("sgd", "binary_crossentropy")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    _, _, _, grads = O.loss_and_grads({k: v.double() for k, v in params.items()}, spec, X.double(), y.double())
    lr = 0.01
    ids, dense = model.split_input(X.to(DEV))
    accum = torch.zeros(1, dtype=torch.float64, device=DEV)
    model.train()
    model.train_step(ids, dense, y.to(DEV), accum)
    assert abs(accum.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    assert abs(model.optim.pop_reg_loss() - float(z["reg_loss"])) <= 2e-5 * abs(float(z["reg_loss"]))
    # SGD: w_new = w - lr * g  (g includes 2*l2*w) for EVERY row of every table
    for k, p in model.state_dict().items():
        expect = params[k].double() - lr * grads[k]
        scale = (lr * grads[k]).abs().max().item()
        assert_close(p, expect, 0, 2e-4 * scale + 1e-7 * params[k].abs().max().item(), "sgd step " + k)


@pytest.mark.parametrize("name", FIT_CASES)
def test_fit_trajectory_matches_reference(name):
    spec, params, z = load_case(name)
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    opt = str(z["optimizer"])
    model.compile(opt, "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    if float(z["lr"]) > 0:
        for g in model.optim.param_groups:
            g["lr"] = float(z["lr"])
    X, y = z["X"], z["y"]
    names = list(model.feature_index.keys())
    xd = {n: X[:, i].copy() for i, n in enumerate(names)}
    hist = model.fit(xd, y.reshape(-1, 1), batch_size=int(z["batch_size"]), epochs=int(z["epochs"]), verbose=0, shuffle=False,
                     validation_data=(dict(xd), y.reshape(-1, 1)))
    # optimizer trajectories amplify rounding (Adam/RMSprop divide by sqrt(v)): 1e-3 relative on the epoch losses
    assert np.allclose(hist.history["loss"], z["history_loss"], rtol=1e-3), (hist.history["loss"], z["history_loss"])
    assert np.allclose(hist.history["val_binary_crossentropy"], z["history_val_bce"], rtol=1e-3)
    assert np.allclose(hist.history["val_auc"], z["history_val_auc"], atol=5e-3)
    pred = model.predict(dict(xd), batch_size=int(z["batch_size"]))
    assert pred.shape == z["pred"].shape and pred.dtype == np.float64
    assert np.allclose(pred, z["pred"], rtol=2e-3, atol=2e-4)
    sd = model.state_dict()
    for k in sd:
        ref = torch.from_numpy(z["final::" + k])
        moved = (ref - params[k]).abs().max().item()
        assert_close(sd[k], ref, 0, 2e-2 * moved + 1e-6, "final weight " + k)


def test_fit_with_torch_optimizer_instance_and_callbacks(tmp_path):
    """Generic path: user-supplied torch optimizer, EarlyStopping / ModelCheckpoint, shuffle, verbose metrics."""
    from deepctr.callbacks import EarlyStopping, ModelCheckpoint
    spec, params, z = load_case("fit_small_sgd")
    model = build_product_model(spec, DEV)
    model.load_state_dict(params, strict=True)
    model.compile(torch.optim.SGD(model.parameters(), lr=0.01), "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
    X, y = z["X"], z["y"]
    names = list(model.feature_index.keys())
    xd = {n: X[:, i].copy() for i, n in enumerate(names)}
    ck = str(tmp_path / "best.pth")
    hist = model.fit(xd, y.reshape(-1, 1), batch_size=32, epochs=2, verbose=2, shuffle=False, validation_split=0.0,
                     validation_data=(dict(xd), y.reshape(-1, 1)),
                     callbacks=[EarlyStopping(monitor="val_auc", patience=5, mode="max"),
                                ModelCheckpoint(ck, monitor="val_auc", save_best_only=True, save_weights_only=True, mode="max")])
    assert np.allclose(hist.history["loss"], z["history_loss"], rtol=1e-3)
    assert set(hist.history) >= {"loss", "binary_crossentropy", "auc", "val_binary_crossentropy", "val_auc"}
    sd = torch.load(ck, map_location="cpu")
    assert set(sd.keys()) == set(params.keys())


def test_no_cpu_fallback():
    spec, params, z = load_case("xdeepfm_small")
    model = build_product_model(spec, "cpu")
    with pytest.raises(RuntimeError):
        model.forward_ids(torch.zeros(2, 5, dtype=torch.int32), torch.zeros(2, 3))


@pytest.mark.parametrize("optimizer", ["adam", "sgd"])
def test_cuda_graph_replay_equals_eager_steps(optimizer):
    """train_step captures itself into a CUDA graph on the third call with the same shapes; replayed steps must leave exactly the
    weights, optimizer state and loss that eager launches leave."""
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    batches = [(X[i * 16:(i + 1) * 16], y[i * 16:(i + 1) * 16]) for i in range(4)] * 2
    out = []
    for graph in (False, True):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile(optimizer, "binary_crossentropy")
        model.use_cuda_graph = graph
        model.train()
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        for Xb, yb in batches:
            ids, dense = model.split_input(Xb.to(DEV))
            model.train_step(ids, dense, yb.to(DEV), accum)
        assert bool(model._graphs) == graph
        reg = model.optim.pop_reg_loss()
        out.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, accum.item(), reg))
    (sd_e, loss_e, reg_e), (sd_g, loss_g, reg_g) = out
    for k in sd_e:
        assert torch.equal(sd_e[k], sd_g[k]), k
    assert abs(loss_e - loss_g) <= 1e-9 * abs(loss_e) and abs(reg_e - reg_g) <= 1e-6 * abs(reg_e)


def test_fit_verbose_metrics_on_the_fused_graph_path():
    """verbose > 0: per-step train metrics come from a device-side prediction log (filled from the CUDA graph's static output);
    losses must equal the verbose = 0 run and the History keys those of the reference."""
    spec, params, z = load_case("fit_small_adam")
    X, y = z["X"], z["y"]
    hists = []
    for verbose in (0, 2):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("adam", "binary_crossentropy", metrics=["binary_crossentropy", "auc"])
        names = list(model.feature_index.keys())
        xd = {n: X[:, i].copy() for i, n in enumerate(names)}
        hists.append(model.fit(xd, y.reshape(-1, 1), batch_size=16, epochs=2, verbose=verbose, shuffle=False,
                               validation_data=(dict(xd), y.reshape(-1, 1))).history)
        assert model._graphs, "6 steps per epoch: the step should have been captured"
    assert np.allclose(hists[0]["loss"], hists[1]["loss"], rtol=1e-6)
    assert set(hists[1]) >= {"loss", "binary_crossentropy", "auc", "val_binary_crossentropy", "val_auc"}
    assert all(0.0 <= v <= 1.0 for v in hists[1]["auc"])


def test_failed_graph_capture_falls_back_to_eager_launches_and_leaves_the_rng_usable():
    """A step that cannot be captured (here: a stream synchronisation injected into the capture) must not poison the process:
    training continues on eager launches with the same results and torch's CUDA generator keeps working (a capture that dies in
    capture_end() otherwise leaves it flagged as capturing and every later randn / model construction raises)."""
    import warnings
    spec, params, z = load_case("xdeepfm_small_zipf")
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"])
    batches = [(X[i * 16:(i + 1) * 16], y[i * 16:(i + 1) * 16]) for i in range(4)]
    out = []
    for sabotage in (False, True):
        model = build_product_model(spec, DEV)
        model.load_state_dict(params, strict=True)
        model.compile("adam", "binary_crossentropy")
        model.use_cuda_graph = sabotage
        if sabotage:
            orig = model._train_step_eager

            def bad(*a, _orig=orig, _m=model, **k):
                if _m._capturing:
                    torch.cuda.current_stream().synchronize()          # illegal while capturing
                return _orig(*a, **k)
            model._train_step_eager = bad
        model.train()
        accum = torch.zeros(1, dtype=torch.float64, device=DEV)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for Xb, yb in batches:
                ids, dense = model.split_input(Xb.to(DEV))
                model.train_step(ids, dense, yb.to(DEV), accum)
        if sabotage:
            assert model._graph_failed and not model._graphs
        out.append(({k: v.detach().clone() for k, v in model.state_dict().items()}, accum.item()))
        assert torch.isfinite(torch.randn(8, device=DEV)).all()          # generator usable after the failed capture
    (sd_a, loss_a), (sd_b, loss_b) = out
    for k in sd_a:
        assert torch.equal(sd_a[k], sd_b[k]), k
    assert abs(loss_a - loss_b) <= 1e-9 * abs(loss_a)
    build_product_model(spec, DEV)                                       # model construction (nn.init on cuda) still works
