"""Fused optimizer kernels vs torch.optim on the CPU (same inputs, several steps, dense-table semantics)."""
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("lazy", [False, True], ids=["dense-pass", "lazy-replay"])
@pytest.mark.parametrize("kind", ["sgd", "adam", "adagrad", "rmsprop"])
def test_fused_optimizer_matches_torch(kind, lazy):
    from deepctr import ops
    from deepctr.optim import FusedOptimizer, TableSet
    g = torch.Generator().manual_seed(5)
    vocab, D, l2_tab, l2_dense = [50, 7], 6, 1e-3, 1e-2
    tabs0 = [torch.randn(v, D, generator=g) for v in vocab]
    dense0 = [torch.randn(5, 3, generator=g), torch.randn(4, generator=g)]
    # reference arm: torch.optim on CPU with dense gradients (grad + 2*l2*w), as the reference trains
    ref_params = [t.clone().requires_grad_(True) for t in tabs0 + dense0]
    ctor = {"sgd": lambda p: torch.optim.SGD(p, lr=0.01), "adam": torch.optim.Adam, "adagrad": torch.optim.Adagrad,
            "rmsprop": torch.optim.RMSprop}[kind]
    ref_opt = ctor(ref_params)
    # product arm
    tabs = [torch.nn.Parameter(t.clone().to(DEV)) for t in tabs0]
    dense = [torch.nn.Parameter(t.clone().to(DEV)) for t in dense0]
    plan = ops.SparsePlan([0, 1], vocab, D)
    plan.sparse_grad = True
    l2map = {id(dense[0]): l2_dense}
    opt = FusedOptimizer(kind, [("w", dense[0]), ("b", dense[1])], [TableSet(plan, tabs, l2_tab)], l2map)
    opt.lazy_tables = lazy
    cache = ops.SegmentCache()
    B = 40
    for step in range(4):
        ids = torch.stack([torch.randint(0, v, (B,), generator=g) for v in vocab], 1).to(torch.int32)
        dout = torch.randn(B, 2, D, generator=g)
        gd = [torch.randn(5, 3, generator=g), torch.randn(4, generator=g)]
        # reference grads
        ref_opt.zero_grad()
        for f in range(2):
            gt = torch.zeros(vocab[f], D).index_add_(0, ids[:, f].long(), dout[:, f])
            ref_params[f].grad = gt + 2 * l2_tab * ref_params[f].detach()
        ref_params[2].grad = gd[0] + 2 * l2_dense * ref_params[2].detach()
        ref_params[3].grad = gd[1].clone()
        ref_opt.step()
        # product
        opt.zero_grad()
        opt.prepare()
        ids_d = ids.to(DEV)
        opt.catch_up(plan, cache, ids_d)          # lazy mode: rows about to be read are replayed first (no-op otherwise)
        out = ops.SparseGather.apply(plan, cache, ids_d, *tabs)
        out.backward(dout.to(DEV))
        dense[0].grad.copy_(gd[0])
        dense[1].grad.copy_(gd[1])
        opt.step(apply_l2=True)
    opt.flush()
    for i, (a, b) in enumerate(zip(tabs + dense, ref_params)):
        moved = (b.detach() - (tabs0 + dense0)[i]).abs().max().item()
        assert_close(a, b.detach(), 0, 1e-4 * moved + 1e-7, "%s param %d" % (kind, i))


def test_sparse_embedding_update_only_touches_batch_rows():
    from deepctr import ops
    from deepctr.optim import FusedOptimizer, TableSet
    tab = torch.nn.Parameter(torch.ones(10, 4, device=DEV))
    w = torch.nn.Parameter(torch.ones(3, device=DEV))
    plan = ops.SparsePlan([0], [10], 4)
    plan.sparse_grad = True
    opt = FusedOptimizer("adam", [("w", w)], [TableSet(plan, [tab], 1e-2)], {})
    opt.sparse_embedding_update = True
    opt.prepare()
    ids = torch.tensor([[2], [7], [2]], dtype=torch.int32, device=DEV)
    ops.SparseGather.apply(plan, ops.SegmentCache(), ids, tab).backward(torch.ones(3, 1, 4, device=DEV))
    opt.step(apply_l2=True)
    changed = (tab.detach().cpu() != 1).any(dim=1)
    assert changed.tolist() == [False, False, True, False, False, False, False, True, False, False]


@pytest.mark.parametrize("kind,hist_cap", [("adam", 1 << 16), ("adam", 4), ("rmsprop", 1 << 16), ("adagrad", 5), ("sgd", 1 << 16)])
def test_lazy_replay_is_bit_identical_to_the_dense_pass(kind, hist_cap):
    """Postponing the updates of untouched rows (catch-up on lookup, flush at the end) must not change a single bit of the tables,
    the optimizer state or the accumulated regulariser compared with streaming every row every step."""
    from deepctr import ops
    from deepctr.optim import FusedOptimizer, TableSet
    g = torch.Generator().manual_seed(11)
    vocab, D, B, steps = [300, 9, 2000], 8, 64, 11
    tabs0 = [torch.randn(v, D, generator=g) for v in vocab]
    lin0 = [torch.randn(v, 1, generator=g) for v in vocab]
    batches = []
    for s in range(steps):
        ids = torch.stack([torch.clamp((float(v) ** torch.rand(B, generator=g)).long() - 1, 0, v - 1) for v in vocab], 1).to(torch.int32)
        batches.append((ids.to(DEV), torch.randn(B, 3, D, generator=g).to(DEV), torch.randn(B, generator=g).to(DEV)))
    results = []
    for lazy in (False, True):
        tabs = [torch.nn.Parameter(t.clone().to(DEV)) for t in tabs0]
        lins = [torch.nn.Parameter(t.clone().to(DEV)) for t in lin0]
        w = torch.nn.Parameter(torch.ones(3, device=DEV))
        plan, plan_lin = ops.SparsePlan([0, 1, 2], vocab, D), ops.SparsePlan([0, 1, 2], vocab, 1)
        plan.sparse_grad = plan_lin.sparse_grad = True
        opt = FusedOptimizer(kind, [("w", w)], [TableSet(plan, tabs, 1e-3), TableSet(plan_lin, lins, 1e-2)], {})
        opt.lazy_tables = lazy
        opt._hist_cap = hist_cap
        cache = ops.SegmentCache()
        for ids, dout, dlin in batches:
            opt.zero_grad()
            opt.prepare()
            cache.clear()
            opt.catch_up(plan, cache, ids)
            opt.catch_up(plan_lin, cache, ids)
            out = ops.SparseGather.apply(plan, cache, ids, *tabs)
            lin = ops.LinearTerm.apply(plan_lin, cache, ids, None, None, *lins)
            torch.autograd.backward([out, lin], [dout, dlin.view(-1, 1)])
            opt.step(apply_l2=True)
        reg = opt.pop_reg_loss()          # flushes
        ts = opt.table_sets
        results.append(([p.detach().clone() for p in tabs + lins], [t.clone() for t in (ts[0].s1 or []) + (ts[1].s1 or [])],
                        [t.clone() for t in (ts[0].s2 or []) + (ts[1].s2 or [])], reg))
    dense_r, lazy_r = results
    for a, b in zip(dense_r[0] + dense_r[1] + dense_r[2], lazy_r[0] + lazy_r[1] + lazy_r[2]):
        assert torch.equal(a, b)
    assert abs(dense_r[3] - lazy_r[3]) <= 1e-6 * abs(dense_r[3])      # same terms, different summation order (float atomics)


def test_packed_replay_is_bit_identical_to_the_scalar_replay():
    """The lazy replay of 4-wide Adam pieces runs on packed fp32 pairs (fma / mul .f32x2); the diagnostic switch turns it off.
    Tables and both moments must not differ in a single bit (this caught ptxas contracting mul.rn.f32x2 + sub.rn.f32x2 into FFMA2)."""
    from deepctr import _native as N, ops
    from deepctr.optim import FusedOptimizer, TableSet
    L = N.lib()

    def run(packed):
        N.check(L.xdfm_set_replay_packed(int(packed)))
        g = torch.Generator().manual_seed(11)
        vocab, D, B = [300, 9, 2000], 8, 64
        tabs = [torch.nn.Parameter(torch.randn(v, D, generator=g).to(DEV)) for v in vocab]
        w = torch.nn.Parameter(torch.ones(3, device=DEV))
        plan = ops.SparsePlan([0, 1, 2], vocab, D)
        plan.sparse_grad = True
        opt = FusedOptimizer("adam", [("w", w)], [TableSet(plan, tabs, 1e-3)], {})
        opt.lazy_tables = True
        cache = ops.SegmentCache()
        for s in range(9):
            ids = torch.stack([torch.clamp((float(v) ** torch.rand(B, generator=g)).long() - 1, 0, v - 1) for v in vocab], 1)
            ids = ids.to(torch.int32).to(DEV)
            dout = torch.randn(B, 3, D, generator=g).to(DEV)
            opt.zero_grad()
            opt.prepare()
            cache.clear()
            opt.catch_up(plan, cache, ids)
            out = ops.SparseGather.apply(plan, cache, ids, *tabs)
            torch.autograd.backward([out], [dout])
            opt.step(apply_l2=True)
        opt.flush()
        ts = opt.table_sets[0]
        return [p.detach().clone() for p in tabs] + [t.clone() for t in ts.s1] + [t.clone() for t in ts.s2]

    try:
        scalar, packed = run(0), run(1)
    finally:
        N.check(L.xdfm_set_replay_packed(1))
    for a, b in zip(scalar, packed):
        assert torch.equal(a, b)
