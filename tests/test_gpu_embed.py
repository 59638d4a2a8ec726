"""GPU parity: input split, fused multi-table gather (bit-exact), linear term, deterministic segmented scatter-add."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tables(vocab, D, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(v, D, generator=g) for v in vocab]


@pytest.mark.parametrize("D", [16, 10, 8, 64, 1, 3])
@pytest.mark.parametrize("B", [1, 257, 4096])
def test_gather_bit_exact(D, B):
    from deepctr import ops
    vocab = [7, 1000, 3, 50000, 29, 2]
    tabs = _tables(vocab, D)
    g = torch.Generator().manual_seed(B)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], 1).to(torch.int32)
    plan = ops.SparsePlan(list(range(len(vocab))), vocab, D)
    out = ops.SparseGather.apply(plan, ops.SegmentCache(), ids.to(DEV), *[t.to(DEV) for t in tabs])
    ref = torch.stack([tabs[f][ids[:, f].long()] for f in range(len(vocab))], 1)
    assert out.shape == (B, len(vocab), D)
    assert torch.equal(out.cpu(), ref)          # index work: bit-exact


def test_gather_shared_table_and_empty_batch():
    from deepctr import ops
    tabs = _tables([11, 5], 4)
    plan = ops.SparsePlan([0, 1, 0], [11, 5], 4)   # features 0 and 2 share table 0
    ids = torch.tensor([[1, 2, 10], [0, 4, 3]], dtype=torch.int32)
    out = ops.SparseGather.apply(plan, ops.SegmentCache(), ids.to(DEV), *[t.to(DEV) for t in tabs]).cpu()
    assert torch.equal(out[:, 0], tabs[0][ids[:, 0].long()]) and torch.equal(out[:, 2], tabs[0][ids[:, 2].long()])
    assert torch.equal(out[:, 1], tabs[1][ids[:, 1].long()])
    empty = ops.SparseGather.apply(plan, ops.SegmentCache(), torch.zeros((0, 3), dtype=torch.int32, device=DEV),
                                   *[t.to(DEV) for t in tabs])
    assert empty.shape == (0, 3, 4)


def test_split_input_truncates_like_long():
    from deepctr import ops
    X = torch.tensor([[3.0, 0.25, 7.9, -0.5], [16777215.0, 1.5, 0.0, 2.0]])
    ids, dense = ops.split_input(X.to(DEV), [0, 2], [1, 3])
    assert torch.equal(ids.cpu(), X[:, [0, 2]].long().to(torch.int32))
    assert torch.equal(dense.cpu(), X[:, [1, 3]])


def test_linear_term_matches():
    from deepctr import ops
    vocab = [13, 700, 5]
    lin = _tables(vocab, 1, seed=3)
    B, nd = 333, 4
    g = torch.Generator().manual_seed(1)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], 1).to(torch.int32)
    dense = torch.rand(B, nd, generator=g)
    w = torch.randn(nd, 1, generator=g)
    plan = ops.SparsePlan([0, 1, 2], vocab, 1)
    out = ops.LinearTerm.apply(plan, ops.SegmentCache(), ids.to(DEV), dense.to(DEV), w.to(DEV), *[t.to(DEV) for t in lin])
    ref = sum(lin[f][ids[:, f].long()] for f in range(3)).double() + dense.double() @ w.double()
    assert_close(out, ref, 1e-6, 1e-6, "linear term")


@pytest.mark.parametrize("D", [16, 10, 64, 4])
@pytest.mark.parametrize("zipf", [False, True])
def test_scatter_add_matches_and_is_deterministic(D, zipf):
    from deepctr import ops
    vocab = [3, 100000, 17, 5000]       # vocab 3 -> runs of ~B/3 entries (long-segment path)
    B = 6000
    g = torch.Generator().manual_seed(7)
    cols = []
    for v in vocab:
        if zipf:
            cols.append(torch.clamp((v ** torch.rand(B, generator=g)).long() - 1, 0, v - 1))
        else:
            cols.append(torch.randint(0, v, (B,), generator=g))
    ids = torch.stack(cols, 1).to(torch.int32)
    tabs = [t.to(DEV).requires_grad_(True) for t in _tables(vocab, D)]
    plan = ops.SparsePlan(list(range(4)), vocab, D)
    dout = torch.randn(B, 4, D, generator=g)
    grads = []
    for rep in range(2):
        for t in tabs:
            t.grad = None
        out = ops.SparseGather.apply(plan, ops.SegmentCache(), ids.to(DEV), *tabs)
        out.backward(dout.to(DEV))
        grads.append([t.grad.clone() for t in tabs])
    for f in range(4):
        ref = torch.zeros(vocab[f], D, dtype=torch.float64).index_add_(0, ids[:, f].long(), dout[:, f].double())
        scale = ref.abs().max().item()
        assert_close(grads[0][f], ref, 1e-5, 1e-6 * scale, "table %d grad" % f)
        assert torch.equal(grads[0][f], grads[1][f]), "scatter-add is not bit-reproducible"


def test_sparse_stash_matches_dense_grad():
    from deepctr import ops
    vocab = [9, 300]
    D, B = 8, 500
    g = torch.Generator().manual_seed(11)
    ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], 1).to(torch.int32).to(DEV)
    tabs = [t.to(DEV).requires_grad_(True) for t in _tables(vocab, D)]
    dout = torch.randn(B, 2, D, generator=g).to(DEV)
    plan = ops.SparsePlan([0, 1], vocab, D)
    ops.SparseGather.apply(plan, ops.SegmentCache(), ids, *tabs).backward(dout)
    dense = torch.cat([t.grad for t in tabs], 0)
    plan2 = ops.SparsePlan([0, 1], vocab, D)
    plan2.sparse_grad = True
    ops.SparseGather.apply(plan2, ops.SegmentCache(), ids, *tabs).backward(dout)
    (uniq, seg_off, pos, nseg, n), gsum = plan2.stash
    k = int(nseg.item())
    keys = uniq[:k].long() & 0xFFFFFFFF
    assert torch.equal(keys, torch.unique(keys)), "unique keys must be sorted and unique"
    rebuilt = torch.zeros_like(dense)
    rebuilt[keys] = gsum[:k]
    assert torch.equal(rebuilt, dense)
