"""Row-sharded tables over peer memory + data-parallel dense part (deepctr/distributed.py, csrc/shard.cu).

Single-GPU coverage: G ranks are EMULATED inside one process -- G ShardedSparse objects wired with raw pointers, their kernels run
one after another on one stream (no kernel waits on another, so this is legal on one GPU) -- plus the world_size-1 NCCL path of
model.distribute().  With >= 2 GPUs the real multi-process test runs as well."""
import os
import tempfile

import numpy as np
import pytest
import torch

from deepctr import _native as N
from deepctr import distributed as D
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _make_ranks(G, rows, table_of, Dm, cap, seed=0):
    g = torch.Generator().manual_seed(seed)
    full_emb = [torch.randn(V, Dm, generator=g) for V in rows]
    full_lin = [torch.randn(V, 1, generator=g) for V in rows]
    vocab = [rows[t] for t in table_of]
    ranks = [D.ShardedSparse(r, G, table_of, rows, vocab, Dm, DEV) for r in range(G)]
    for r, sh in enumerate(ranks):
        for t, V in enumerate(rows):
            a, k = sh.base[r][t], D.shard_rows(V, r, G)
            sh.emb[a:a + k].copy_(D.take_shard(full_emb[t], r, G))
            sh.lin[a:a + k].copy_(D.take_shard(full_lin[t], r, G))
        sh.alloc_exchange(cap)
    infos = [sh.local_pointers() for sh in ranks]
    for sh in ranks:
        sh.connect(infos)
    return ranks, full_emb, full_lin


@pytest.mark.parametrize("G", [1, 2, 3])
def test_sharded_gather_is_bit_exact(G):
    rows, table_of, Dm, B = [50, 7, 1000, 3], [0, 1, 2, 3, 2], 16, 333
    ranks, full_emb, full_lin = _make_ranks(G, rows, table_of, Dm, B * len(table_of))
    g = torch.Generator().manual_seed(1)
    ids = torch.stack([torch.randint(0, rows[t], (B,), generator=g) for t in table_of], 1).to(torch.int32)
    dense, w = torch.rand(B, 3, generator=g), torch.randn(3, 1, generator=g)
    ref = torch.stack([full_emb[t][ids[:, f].long()] for f, t in enumerate(table_of)], 1)
    ref_lin = sum(full_lin[t][ids[:, f].long(), 0] for f, t in enumerate(table_of))
    for sh in ranks:           # every rank sees the same (peer-mapped) tables
        out, lin = sh.gather(ids.to(DEV), want_emb=True, dense=dense.to(DEV), dense_w=w.to(DEV), want_lin=True)
        assert torch.equal(out.cpu(), ref)
        assert_close(lin, ref_lin + (dense @ w).reshape(-1), 1e-6, 1e-6, "first-order term")
    out, _ = ranks[0].gather(ids[:0].to(DEV))
    assert out.shape == (0, len(table_of), Dm)


@pytest.mark.parametrize("G", [1, 3])
def test_lookup_through_distinct_rows_equals_the_direct_lookup(G):
    """The default forward fetches every DISTINCT row of the batch once and expands locally; it must return the bits of the direct
    per-lookup kernels, serve the embedding and the first-order call of one forward from one fetch, notice a changed batch or
    changed tables, and hand its segments to the backward."""
    rows, table_of, Dm, B = [50, 7, 1000, 3, 4000], [0, 1, 2, 3, 4, 2], 16, 513
    m = len(table_of)
    ranks, full_emb, full_lin = _make_ranks(G, rows, table_of, Dm, B * m)
    sh = ranks[G - 1]
    g = torch.Generator().manual_seed(5)

    def batch():
        cols = [torch.clamp((rows[t] ** torch.rand(B, generator=g)).long() - 1, 0, rows[t] - 1) for t in table_of]
        return torch.stack(cols, 1).to(torch.int32).to(DEV)

    ids = batch()
    dense, w = torch.rand(B, 3, generator=g).to(DEV), torch.randn(3, 1, generator=g).to(DEV)
    sh.unique_lookup = False
    out0, _ = sh.gather(ids, want_emb=True)
    _, lin0 = sh.gather(ids, want_emb=False, dense=dense, dense_w=w, want_lin=True)
    sh.unique_lookup = True
    out1, _ = sh.gather(ids, want_emb=True)
    assert sh._uniq == {"emb"}
    _, lin1 = sh.gather(ids, want_emb=False, dense=dense, dense_w=w, want_lin=True)
    assert sh._uniq == {"emb", "lin"}                      # second call of the forward: expanded from the same fetch
    assert torch.equal(out0, out1) and torch.equal(lin0, lin1)
    nseg = int(sh.nseg.item())
    assert nseg == sum(int(ids[:, [f for f, t in enumerate(table_of) if t == tt]].unique().numel()) for tt in range(len(rows)))
    # the backward takes the forward's segments over (no second sort) and gives the sums of the direct design
    demb, dlin = torch.randn(B, m, Dm, generator=g).to(DEV), torch.randn(B, generator=g).to(DEV)
    sh.stash = {"ids": ids, "demb": demb, "dlin": dlin}
    assert sh._segments_current(ids)
    sh.reduce_local()
    keys1, gs1, gl1 = sh.x_keys[:nseg].clone(), sh.x_gsum[:nseg].clone(), sh.x_gsum_lin[:nseg].clone()
    sh.unique_lookup = False
    sh.stash = {"ids": ids, "demb": demb, "dlin": dlin}
    sh.reduce_local()
    assert torch.equal(keys1, sh.x_keys[:nseg]) and torch.equal(gs1, sh.x_gsum[:nseg]) and torch.equal(gl1, sh.x_gsum_lin[:nseg])
    # a rewritten table row and a batch changed in place are both seen by the next forward
    sh.unique_lookup = True
    sh.gather(ids, want_emb=True)
    sh.emb.mul_(2.0)
    sh.tables_changed()
    ids.copy_(batch())
    out2, _ = sh.gather(ids, want_emb=True)
    sh.unique_lookup = False
    out3, _ = sh.gather(ids, want_emb=True)
    assert torch.equal(out2, out3)


@pytest.mark.parametrize("G,zipf", [(1, False), (2, True), (4, True)])
def test_sharded_backward_exchange_matches_scatter_add_and_is_deterministic(G, zipf):
    rows, table_of, Dm, B = [50, 7, 1000, 3, 29], [0, 1, 2, 3, 4, 2], 8, 257
    m = len(table_of)
    ranks, full_emb, full_lin = _make_ranks(G, rows, table_of, Dm, B * m)
    g = torch.Generator().manual_seed(2)
    row_off = np.concatenate([[0], np.cumsum(rows)])
    acc = torch.zeros(int(row_off[-1]), Dm, dtype=torch.float64)
    acc_lin = torch.zeros(int(row_off[-1]), dtype=torch.float64)
    for r, sh in enumerate(ranks):      # every emulated rank posts its own batch
        Br = B - 17 * r
        cols = []
        for t in table_of:
            u = torch.rand(Br, generator=g)
            cols.append(torch.clamp((rows[t] ** u).long() - 1, 0, rows[t] - 1) if zipf else torch.randint(0, rows[t], (Br,), generator=g))
        ids = torch.stack(cols, 1).to(torch.int32)
        demb, dlin = torch.randn(Br, m, Dm, generator=g), torch.randn(Br, generator=g)
        for f, t in enumerate(table_of):
            acc.index_add_(0, ids[:, f].long() + int(row_off[t]), demb[:, f].double())
            acc_lin.index_add_(0, ids[:, f].long() + int(row_off[t]), dlin.double())
        sh.stash = {"ids": ids.to(DEV), "demb": demb.to(DEV).contiguous(), "dlin": dlin.to(DEV)}
        sh.reduce_local()
    results = []
    for rep in range(2):
        per_rank = []
        for r, sh in enumerate(ranks):      # owners pull and merge
            sh.pull_segments()
            n = int(sh.p_nseg.item())
            per_rank.append((sh.p_uniq[:n].clone(), sh.p_gsum[:n].clone(), sh.p_gsum_lin[:n].clone()))
        results.append(per_rank)
    touched = 0
    for r, sh in enumerate(ranks):
        uniq, gs, gl = results[0][r]
        assert torch.equal(uniq, results[1][r][0]) and torch.equal(gs, results[1][r][1]) and torch.equal(gl, results[1][r][2])
        uniq = uniq.cpu().long()
        assert bool((uniq[1:] > uniq[:-1]).all())
        # local row -> global row
        glob = torch.empty_like(uniq)
        for t, V in enumerate(rows):
            a, k = sh.base[r][t], D.shard_rows(V, r, G)
            sel = (uniq >= a) & (uniq < a + k)
            glob[sel] = (uniq[sel] - a) * G + r + int(row_off[t])
        assert_close(gs, acc[glob], 1e-5, 1e-5, "row sums of rank %d" % r)
        assert_close(gl, acc_lin[glob], 1e-5, 1e-5, "first-order sums of rank %d" % r)
        touched += uniq.numel()
    assert touched == int((acc.abs().sum(1) > 0).sum())


def _run_in_subprocess(world, optimizer):
    import torch.multiprocessing as mp
    from tests import dist_worker
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(dist_worker.spawn_entry, args=(world, os.path.join(d, "init"), optimizer), nprocs=world, join=True)


def test_deferred_tables_world_size_1():
    _run_in_subprocess(1, "deferred")


@pytest.mark.parametrize("optimizer", ["adam", "sgd"])
def test_distribute_world_size_1_equals_plain_model(optimizer):
    _run_in_subprocess(1, optimizer)


def test_pro_distribute_world_size_1_equals_plain_model():
    _run_in_subprocess(1, "pro:adam")


def test_multi_value_features_distribute_world_size_1_equals_plain_model():
    """VarLenSparseFeat columns over row-sharded tables: slots looked up through the batch's distinct rows, pooled per field."""
    _run_in_subprocess(1, "varlen:adam")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("optimizer", ["adam", "adagrad", "pro:adam", "varlen:adam"])
def test_two_gpus_equal_one_gpu(optimizer):
    _run_in_subprocess(2, optimizer)
