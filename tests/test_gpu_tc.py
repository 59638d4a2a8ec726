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
