"""Multi-process worker of the distributed parity test (also runnable under torchrun on an N-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py

Every rank trains its slice of the same global batches with the hybrid-parallel model (row-sharded tables over NVLink peer
memory + all-reduced dense gradients); rank 0 also trains an ordinary single-GPU model on the concatenated batches and the
two must agree: per-step losses, predictions, and every tensor of the (re-assembled) state_dict."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "xdeepfm-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import xdeepfm_oracle as O  # noqa: E402
from tests.helpers import assert_close, build_product_model  # noqa: E402


def spec_for_test(variant="xdeepfm"):
    if variant == "pro":
        # xDeepFM Pro (SFG decoder, label-aware attention): BASELINE configs[3] is its data-parallel run
        return O.ModelSpec(sparse_names=["C%d" % i for i in range(1, 7)], vocab_sizes=[50, 7, 1000, 3, 29, 400], embedding_dim=16,
                           dense_names=["I1", "I2", "I3"], cin_layer_size=(32, 16), dnn_hidden_units=(64, 32), l2_reg_linear=1e-4,
                           l2_reg_embedding=1e-4, l2_reg_dnn=1e-4, l2_reg_cin=1e-4, variant="pro", sfg_hidden_units=(32, 16))
    return O.ModelSpec(sparse_names=["C%d" % i for i in range(1, 7)], vocab_sizes=[50, 7, 1000, 3, 29, 400], embedding_dim=16,
                       dense_names=["I1", "I2", "I3"], cin_layer_size=(32, 16), dnn_hidden_units=(64, 32), l2_reg_linear=1e-4,
                       l2_reg_embedding=1e-4, l2_reg_dnn=1e-4, l2_reg_cin=1e-4)


def _setup(variant, dev):
    """(model builder, initial state_dict, make_inputs(n, seed, zipf) -> (X float [n, columns], y [n]))."""
    if variant == "varlen":
        # multi-value features (VarLenSparseFeat: mean / max / sum pooling, id-0 masks and length columns) over row-sharded tables:
        # rows drawn with replacement from the reference-generated fixture, weights = the fixture's
        from tests.helpers import build_varlen_product_model, load_varlen_case
        desc, _, params, z = load_varlen_case("xdeepfm_varlen")
        Xz, yz = torch.from_numpy(z["X"]).float(), torch.from_numpy(z["y"]).float()

        def make_inputs(n, seed, zipf=False):
            idx = torch.randint(0, Xz.shape[0], (n,), generator=torch.Generator().manual_seed(seed))
            return Xz[idx].clone(), yz[idx].clone()

        return (lambda: build_varlen_product_model(desc, dev)), params, make_inputs
    spec = spec_for_test(variant)
    return (lambda: build_product_model(spec, dev)), O.make_params(spec, seed=11), (lambda n, seed, zipf=False: O.make_inputs(spec, n, seed=seed, zipf=zipf))


def _as_dict(model, X):
    return {n: (X[:, a:b].numpy().copy() if b - a > 1 else X[:, a].numpy().copy()) for n, (a, b) in model.feature_index.items()}


def run(rank, world, optimizer="adam", steps=6, per_rank=48, fit_check=True, variant="xdeepfm"):
    from deepctr.distributed import rank_slice
    dev = "cuda:%d" % torch.cuda.current_device()
    build, params, make_inputs = _setup(variant, dev)
    gb = per_rank * world
    batches = [make_inputs(gb, 100 + s, zipf=(s % 2 == 0)) for s in range(steps)]

    model = build()
    model.load_state_dict(params, strict=True)
    model.distribute(max_batch=per_rank)
    model.compile(optimizer, "binary_crossentropy")
    model.train()
    accum = torch.zeros(1, dtype=torch.float64, device=dev)
    losses = []
    for X, y in batches:
        a, b = rank_slice(0, gb, rank, world)
        ids, dense = model.split_input(X[a:b].to(dev))
        accum.zero_()
        model.train_step(ids, dense, y[a:b].to(dev), accum)
        t = accum.clone()
        dist.all_reduce(t)
        losses.append(t.item())
    reg = model.optim.pop_reg_loss()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}          # collective
    model.eval()
    Xe, _ = make_inputs(64, 999)
    with torch.no_grad():
        pred = model(Xe.to(dev)).cpu()

    # checkpoint round trip through the sharded tables (collective)
    model.load_state_dict(sd, strict=True)
    sd2 = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), "state_dict round trip " + k

    if fit_check:
        # fit(): every rank passes the same arrays; the epoch loss must equal the single-GPU value at batch = world * bs
        Xf, yf = make_inputs(5 * gb + 7, 77)
        xd = _as_dict(model, Xf)
        hist = model.fit(xd, yf.numpy().reshape(-1, 1), batch_size=per_rank, epochs=2, verbose=0, shuffle=False)
        fit_loss = list(hist.history["loss"])
    if rank != 0:
        return
    ref = build()
    ref.load_state_dict(params, strict=True)
    ref.compile(optimizer, "binary_crossentropy")
    ref.train()
    ref_losses = []
    for X, y in batches:
        ids, dense = ref.split_input(X.to(dev))
        accum.zero_()
        ref.train_step(ids, dense, y.to(dev), accum)
        ref_losses.append(accum.item())
    ref_reg = ref.optim.pop_reg_loss()
    assert np.allclose(losses, ref_losses, rtol=2e-5), (losses, ref_losses)
    assert abs(reg - ref_reg) <= 1e-5 * abs(ref_reg), (reg, ref_reg)
    rsd = ref.state_dict()
    assert set(rsd.keys()) == set(sd.keys())
    for k in rsd:
        moved = (rsd[k].cpu() - params[k]).abs().max().item()
        assert_close(sd[k], rsd[k], 0, 2e-3 * moved + 1e-7, "weights after %d steps: %s" % (steps, k))
    ref.eval()
    with torch.no_grad():
        assert_close(pred, ref(Xe.to(dev)), 1e-4, 1e-6, "predictions")
    if fit_check:
        hist = ref.fit(xd, yf.numpy().reshape(-1, 1), batch_size=per_rank * world, epochs=2, verbose=0, shuffle=False)
        assert np.allclose(fit_loss, hist.history["loss"], rtol=1e-4), (fit_loss, hist.history["loss"])
    print("dist parity ok: %s world=%d optimizer=%s losses=%s" % (variant, world, optimizer, ["%.4f" % l for l in losses]), flush=True)


def run_deferred(rank, world):
    """Tables that are never materialised whole (deepctr.inputs.deferred_tables): shards are initialised in place; training runs;
    the per-rank checkpoint carries the shard and its routing."""
    from deepctr.distributed import shard_rows, sharded_state_dict
    from deepctr.inputs import deferred_tables
    dev = "cuda:%d" % torch.cuda.current_device()
    spec = spec_for_test()
    with deferred_tables():
        model = build_product_model(spec, dev)
    assert all(e.weight.shape[0] == 1 for e in model.embedding_dict.values())
    model.distribute(max_batch=32)
    model.compile("adam", "binary_crossentropy")
    model.optim.sparse_embedding_update = True
    model.train()
    accum = torch.zeros(1, dtype=torch.float64, device=dev)
    for s in range(4):
        X, y = O.make_inputs(spec, 32, seed=500 + s + 10 * rank)
        ids, dense = model.split_input(X.to(dev))
        model.train_step(ids, dense, y.to(dev), accum)
    assert torch.isfinite(accum).all() and accum.item() > 0
    ck = sharded_state_dict(model)
    sh = model._dist.sharded
    assert ck["emb_shard"].shape == (sum(shard_rows(V, rank, world) for V in spec.vocab_sizes), spec.embedding_dim)
    std = ck["emb_shard"].std().item()
    assert 0.2e-4 < std < 5e-3, std            # N(0, 1e-4) init, a few rows moved by Adam
    assert float(ck["emb_shard"].abs().max()) > 0
    if rank == 0:
        print("deferred tables ok: world=%d local rows=%d" % (world, sh.local_rows), flush=True)


def spawn_entry(rank, world, init_file, optimizer):
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("nccl", init_method="file://" + init_file, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank % torch.cuda.device_count()))
    try:
        if optimizer == "deferred":
            run_deferred(rank, world)
        elif ":" in optimizer:
            run(rank, world, optimizer.split(":")[1], variant=optimizer.split(":")[0])
        else:
            run(rank, world, optimizer)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    try:
        for opt in (sys.argv[1:] or ["adam", "sgd"]):
            if opt == "deferred":
                run_deferred(rank, world)
            elif ":" in opt:
                run(rank, world, opt.split(":")[1], variant=opt.split(":")[0])
            else:
                run(rank, world, opt)
    finally:
        dist.destroy_process_group()
