"""Parity at BASELINE.json's full config-2 sizes (batch 8192, 26 fields, D = 16, CIN (200, 200, 200)) through properties that do
not need a CPU restatement of the whole batch: exact integer arithmetic for the gather / scatter-add, sample independence
(rows of the full batch == the same rows run as a small batch, which the oracle can check) and additivity of the weight
gradient over batch chunks for the tensor-core CIN.  Runs last (file name) because it is the heaviest file."""
import pytest
import torch

from oracle import xdeepfm_oracle as O
from tests.helpers import assert_close
from tests.test_gpu_cin import _rand_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CRITEO_VOCAB = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306, 10,
                5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


def test_full_size_gather_bit_exact_and_scatter_add_exact_on_integer_gradients():
    """B = 8192 x 26 fields x D = 16 (vocabularies capped at 2e5 rows so the CPU side stays small).  Gather: bit-exact.  Backward
    with integer-valued gradients: every per-row sum is exactly representable, so the sorted segmented scatter-add must equal
    index_add_ bit for bit whatever order it sums in, and the grand total must equal the total of the incoming gradient."""
    from deepctr import ops
    B, D = 8192, 16
    vocab = [min(v, 200000) for v in CRITEO_VOCAB]
    g = torch.Generator().manual_seed(2025)
    tabs = [torch.randn(v, D, generator=g) for v in vocab]
    cols = []
    for f, v in enumerate(vocab):
        if f % 2:       # Zipf-like: heavy duplicates (long segments)
            cols.append(torch.clamp((float(v) ** torch.rand(B, generator=g)).long() - 1, 0, v - 1))
        else:
            cols.append(torch.randint(0, v, (B,), generator=g))
    ids = torch.stack(cols, 1).to(torch.int32)
    plan = ops.SparsePlan(list(range(len(vocab))), vocab, D)
    dev_tabs = [t.to(DEV).requires_grad_(True) for t in tabs]
    out = ops.SparseGather.apply(plan, ops.SegmentCache(), ids.to(DEV), *dev_tabs)
    ref = torch.stack([tabs[f][ids[:, f].long()] for f in range(len(vocab))], 1)
    assert torch.equal(out.detach().cpu(), ref)
    dout = torch.randint(-2, 3, (B, len(vocab), D), generator=g).float()
    out.backward(dout.to(DEV))
    total = 0.0
    for f, v in enumerate(vocab):
        want = torch.zeros(v, D, dtype=torch.float64).index_add_(0, ids[:, f].long(), dout[:, f].double())
        got = dev_tabs[f].grad.cpu().double()
        assert torch.equal(got, want), "table %d" % f
        total += got.sum().item()
    assert total == dout.double().sum().item()


def test_full_size_tensor_core_cin_rows_are_batch_independent_and_match_the_oracle():
    """bf16 tensor-core CIN at the full config-2 size.  (1) Sample independence: rows [lo, lo + 40) of the full batch -- pooled
    outputs and dX0 -- equal the same 40 samples run on their own (other tile positions / CTA assignment) within the bf16 tolerance.
    (2) Those rows match the fp64 oracle at the stated bf16 tolerance (2e-2 of the tensor's scale).  (3) Additivity: every weight /
    bias gradient of the full batch equals the sum over four 2048-sample chunks."""
    from deepctr import ops
    B, m, D, layers = 8192, 26, 16, (200, 200, 200)
    x0, Ws, bs, g = _rand_case(B, m, D, layers, True, seed=2025)
    x0 = x0.to(torch.bfloat16).float()
    Ws = [W.to(torch.bfloat16).float() for W in Ws]
    fm = sum(layers[:-1]) // 2 + layers[-1]
    gout = torch.randn(B, fm, generator=g)
    cfg = ops.CINConfig(m, list(layers), True, "relu", pool=True, impl="bf16")

    def run(lo, hi):
        x = x0[lo:hi].to(DEV).requires_grad_(True)
        wb = []
        for W, b in zip(Ws, bs):
            wb += [W.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
        out = ops.cin_apply(cfg, x, *wb)
        out.backward(gout[lo:hi].to(DEV))
        return out.detach(), x.grad, [t.grad for t in wb]

    out, dx, dwb = run(0, B)
    assert torch.isfinite(out).all() and torch.isfinite(dx).all()
    o_scale, dx_scale = out.abs().max().item(), dx.abs().max().item()
    for lo in (0, 4000, B - 40):
        o_s, dx_s, _ = run(lo, lo + 40)
        # identical when both batch sizes take the same kernel; otherwise fp32 accumulation order may flip single bf16 roundings /
        # ReLU decisions on the kink: the stated bf16 tolerance bounds it, a wrong tile or row mapping would be O(1) off
        assert_close(out[lo:lo + 40], o_s, 0, 2e-2 * o_scale, "pooled rows %d.." % lo)
        assert_close(dx[lo:lo + 40], dx_s, 0, 2e-2 * dx_scale, "dX0 rows %d.." % lo)
        want = O.cin_forward(x0[lo:lo + 40].double(), [W.double() for W in Ws], [b.double() for b in bs], True, "relu", pool=True)
        assert_close(o_s, want, 2e-2, 2e-2 * max(want.abs().max().item(), 1.0), "pooled rows %d.. vs oracle" % lo)
    sums = None
    for lo in range(0, B, 2048):
        _, _, part = run(lo, lo + 2048)
        sums = [p.clone() for p in part] if sums is None else [a + p for a, p in zip(sums, part)]
    for i, (full, acc) in enumerate(zip(dwb, sums)):
        assert_close(full, acc, 0, 2e-3 * max(acc.abs().max().item(), 1e-6), "weight / bias gradient %d: full batch vs sum of chunks" % i)
