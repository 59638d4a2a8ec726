"""tcgen05 building-block self-test and the bf16 tensor-core CIN path."""
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 64), (208, 128), (256, 256), (64, 192)])
def test_tcgen05_selftest_gemm(mode, N, K):
    from deepctr import _native as Nv
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    Nv.check(Nv.lib().xdfm_tc_selftest_gemm(Nv.ptr(A), Nv.ptr(B), N, K, mode, Nv.ptr(out), Nv.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float().double() @ B.float().double().t()
    assert_close(out, ref, 1e-5, 1e-4, "selftest mode %d" % mode)


def _tc_layer(x0b, xkb, xk_stride, W, b, m, Hp, H, D, act, direct_begin, fm, col_off, pool=True):
    """One CIN layer through the C ABI (bf16 tensor-core path)."""
    from deepctr import _native as Nv
    L = Nv.lib()
    B = x0b.shape[0]
    n = L.xdfm_cin_tc_wprime_elems(m, Hp, H, D)
    assert n > 0, L.xdfm_last_error()
    wprime = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    yb = torch.full((B, H, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    pooled = torch.full((B, fm), float("nan"), device=DEV)
    maps = torch.full((B, fm, D), float("nan"), device=DEV) if not pool else None
    Nv.check(L.xdfm_cin_fwd_tc(Nv.ptr(x0b), Nv.ptr(xkb), xk_stride, Nv.ptr(W), Nv.ptr(b), Nv.ptr(wprime), B, m, Hp, H, D,
                               Nv.ACT[act], Nv.ptr(yb), direct_begin, Nv.ptr(pooled), Nv.ptr(maps), fm, col_off, Nv.stream_ptr()))
    torch.cuda.synchronize()
    return yb, pooled, maps


def _emulated_layer(x0b, xkb, W, b, act):
    """fp64 accumulation over bf16-rounded operands and bf16-rounded products (what the kernel computes)."""
    Bn, m, D = x0b.shape
    z = (xkb.float()[:, :, None, :] * x0b.float()[:, None, :, :]).to(torch.bfloat16)        # [B, Hp, m, D], k = i*m + j
    z = z.reshape(Bn, -1, D).double()
    Wb = W.reshape(W.shape[0], -1).to(torch.bfloat16).double()
    y = torch.einsum("hk,bkd->bhd", Wb, z) + b.double().view(1, -1, 1)
    return torch.relu(y) if act == "relu" else y


TC_CASES = [
    # B, m, D, H, Hp(=m for the first layer)
    (8, 2, 16, 16, 2),
    (8, 26, 16, 200, 26),
    (19, 26, 16, 200, 100),
    (300, 26, 16, 200, 100),
    (64, 26, 8, 256, 26),
    (33, 26, 8, 128, 128),
    (20, 22, 32, 256, 128),
    (9, 26, 64, 256, 26),
    (5, 26, 128, 64, 26),
    (40, 7, 16, 48, 24),
    (2500, 26, 16, 200, 26),      # 313 tiles -> several tiles per persistent CTA (ring / phase wrap-around)
    (1300, 12, 32, 64, 12),
    (700, 10, 8, 32, 10),
]


@pytest.mark.parametrize("case", TC_CASES, ids=[str(c) for c in TC_CASES])
def test_cin_tc_single_layer_matches_emulation(case):
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case))
    x0 = (torch.randn(B, m, D, generator=g) * 0.5).to(torch.bfloat16)
    Hprev = max(Hp, 2 * Hp if Hp != m else Hp)          # xk is a channel-slice of a wider previous layer when Hp != m
    xk_full = x0 if Hp == m else (torch.randn(B, Hprev, D, generator=g) * 0.5).to(torch.bfloat16)
    xk = xk_full[:, :Hp]
    K = Hp * m
    W = torch.randn(H, K, generator=g) / K ** 0.5
    b = torch.randn(H, generator=g) * 0.1
    hdb = H // 2
    fm = H - hdb + 3
    x0d, xkd, Wd, bd = x0.to(DEV), xk_full.to(DEV), W.to(DEV), b.to(DEV)
    yb, pooled, _ = _tc_layer(x0d, xkd, xk_full.shape[1] * D, Wd, bd, m, Hp, H, D, "relu", hdb, fm, 3)
    ref = _emulated_layer(x0, xk, W, b, "relu")
    scale = ref.abs().max().item()
    # bf16 storage of y: 2^-9 relative; accumulation order differences are ~1e-6
    assert_close(yb.float(), ref, 6e-3, 1e-3 * scale, "y (bf16)")
    assert_close(pooled[:, 3:], ref[:, hdb:].sum(-1), 2e-4, 2e-4 * scale * D ** 0.5, "pooled (fp32 from accumulators)")
