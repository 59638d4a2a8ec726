"""Host-side logic of the multi-GPU path, on CPU: integer routing of the row-sharded tables, and the process-group
semantics (world_size-2 gloo): gradients are SUMMED over ranks, the batch slicing of fit() partitions every global batch."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepctr import distributed as D
from oracle import xdeepfm_oracle as O
from tests.helpers import build_product_model


@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_row_routing_is_a_bijection(G):
    rows = [1, 7, 16, 1000, 3, 29]
    base, total = D.shard_layout(rows, G)
    assert sum(total) == sum(rows)
    S = D.key_stride(total, G)
    assert S > max(total) and S & (S - 1) == 0
    seen = set()
    for t, V in enumerate(rows):
        assert sum(D.shard_rows(V, g, G) for g in range(G)) == V
        for r in range(V):
            g, lr = D.owner_of(r, G), D.local_row(r, G)
            assert 0 <= lr < D.shard_rows(V, g, G)
            slot = (g, base[g][t] + lr)
            assert slot not in seen and slot[1] < total[g]
            seen.add(slot)
            k = D.shard_key(t, r, base, G, S)
            assert k // S == g and k % S == slot[1] and k < (1 << 32)
    assert len(seen) == sum(rows)


def test_key_stride_rejects_overflow():
    with pytest.raises(ValueError):
        D.key_stride([1 << 30], 8)


@pytest.mark.parametrize("G", [1, 2, 5])
def test_take_and_merge_shards_roundtrip(G):
    for V in (1, 4, 11):
        full = torch.arange(V * 3, dtype=torch.float32).reshape(V, 3)
        kmax = D.shard_rows(V, 0, G)
        shards = []
        for g in range(G):
            s = torch.zeros(kmax, 3)
            part = D.take_shard(full, g, G)
            assert part.shape[0] == D.shard_rows(V, g, G)
            s[:part.shape[0]] = part
            shards.append(s)
        assert torch.equal(D.merge_shards(shards, V), full)


def test_rank_slice_partitions_every_batch():
    for n in (0, 1, 5, 8, 13):
        for G in (1, 2, 3, 8):
            cuts = [D.rank_slice(10, 10 + n, r, G) for r in range(G)]
            assert cuts[0][0] == 10 and cuts[-1][1] == 10 + n
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------------------------------
def _gloo_worker(rank, world, init_file, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "xdeepfm-pytorch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        ctx = D.DistContext()
        assert (ctx.rank, ctx.world) == (rank, world)
        # plumbing
        assert ctx.all_gather_object({"r": rank}) == [{"r": r} for r in range(world)]
        t = torch.full((3,), float(rank + 1))
        assert torch.equal(ctx.all_reduce_sum(t), torch.full((3,), float(sum(range(1, world + 1)))))
        b = torch.full((2,), float(rank))
        assert torch.equal(ctx.broadcast(b, 0), torch.zeros(2))
        ctx.barrier("cpu")

        # data-parallel semantics: sum of the ranks' gradients of the SUM-reduced loss == gradient on the concatenated batch,
        # and the L2 term is added once (after the reduction)
        spec = O.ModelSpec(sparse_names=["C1", "C2", "C3"], vocab_sizes=[11, 5, 40], embedding_dim=4, dense_names=["I1", "I2"],
                           cin_layer_size=(8, 4), dnn_hidden_units=(8, 8), l2_reg_linear=1e-3, l2_reg_embedding=1e-3, l2_reg_dnn=1e-3,
                           l2_reg_cin=1e-3)
        params = {k: v.double() for k, v in O.make_params(spec, seed=5).items()}
        X, y = O.make_inputs(spec, 12, seed=5)
        X, y = X.double(), y.double()
        a, e = D.rank_slice(0, 12, rank, world)
        p_loc = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        yp = O.xdeepfm_forward(p_loc, spec, X[a:e]).reshape(-1)
        loss = torch.nn.functional.binary_cross_entropy(yp, y[a:e], reduction="sum")
        loss.backward()
        flat = torch.cat([p_loc[k].grad.reshape(-1) for k in sorted(p_loc)])
        ctx.all_reduce_sum(flat)
        p_reg = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        O.reg_loss(p_reg, spec).backward()
        flat = flat + torch.cat([(torch.zeros_like(p_reg[k]) if p_reg[k].grad is None else p_reg[k].grad).reshape(-1)
                                 for k in sorted(p_reg)])
        _, _, _, grads = O.loss_and_grads(params, spec, X, y)
        ref = torch.cat([grads[k].reshape(-1) for k in sorted(grads)])
        assert torch.allclose(flat, ref, rtol=1e-9, atol=1e-12)
        lsum = ctx.all_reduce_sum(loss.detach().reshape(1).clone())
        full = torch.nn.functional.binary_cross_entropy(O.xdeepfm_forward(params, spec, X).reshape(-1), y, reduction="sum")
        assert abs(lsum.item() - full.item()) < 1e-9

        # fit() batch slicing: same permutation everywhere, every global batch split into per-rank contiguous slices
        model = build_product_model(spec, "cpu")
        model._dist = ctx
        n, bs = 23, 4
        order = torch.randperm(n)                      # differs per rank on purpose: rank 0's is broadcast
        mine = model._dist_local_order(order, n, bs)
        alls = ctx.all_gather_object(mine.tolist())
        order0 = ctx.all_gather_object(order.tolist())[0]
        got = []
        steps = (n - 1) // (bs * world) + 1
        for s in range(steps):
            for r in range(world):
                lo = s * bs
                got += alls[r][lo:lo + bs] if s < steps - 1 else alls[r][lo:]
        assert got == order0
        assert sorted(sum(alls, [])) == list(range(n))
        seq = model._dist_local_order(None, n, bs)
        assert sorted(sum(ctx.all_gather_object(seq.tolist()), [])) == list(range(n))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_semantics():
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_gloo_worker, args=(2, init_file, d), nprocs=2, join=True)
        assert os.path.exists(os.path.join(d, "ok0")) and os.path.exists(os.path.join(d, "ok1"))
