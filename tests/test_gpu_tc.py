"""tcgen05 building-block self-test and the bf16 tensor-core CIN forward kernel (through the C ABI)."""
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


@pytest.mark.parametrize("N,K", [(112, 208), (32, 48), (16, 16), (128, 144), (256, 96), (112, 192)])
def test_tcgen05_selftest_gemm_k_tail_in_32_byte_swizzle(N, K):
    """Mode 2: A in TMEM, B = full 64-wide SWIZZLE_128B chunks + 16-wide SWIZZLE_32B boxes for the K tail (the layout the dX kernel
    streams W'' in when H is not a multiple of 64)."""
    from deepctr import _native as Nv
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    Nv.check(Nv.lib().xdfm_tc_selftest_gemm(Nv.ptr(A), Nv.ptr(B), N, K, 2, Nv.ptr(out), Nv.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float().double() @ B.float().double().t()
    assert_close(out, ref, 1e-5, 1e-4, "selftest mode 2")


@pytest.mark.parametrize("N,K", [(64, 64), (128, 128), (208, 64), (208, 128), (16, 16), (256, 256), (112, 48)])
def test_tcgen05_selftest_gemm_mn_major_b(N, K):
    """Mode 3: B handed over transposed ([K, N] row-major) and read MN-major by the tensor core -- how the weight-gradient kernels
    consume row-layout activations without a transposed copy."""
    from deepctr import _native as Nv
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    Bt = B.t().contiguous().to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    Nv.check(Nv.lib().xdfm_tc_selftest_gemm(Nv.ptr(A), Nv.ptr(Bt), N, K, 3, Nv.ptr(out), Nv.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float().double().cpu() @ B.float().double().t()
    assert_close(out, ref, 1e-5, 1e-4, "selftest mode 3")


def to_rows(x, CP):
    """[B, C, D] fp32 (device) -> row layout [B*D, CP] bf16 via the library."""
    from deepctr import _native as Nv
    B, C, D = x.shape
    xt = torch.empty((B * D, CP), dtype=torch.bfloat16, device=DEV)
    Nv.check(Nv.lib().xdfm_to_rows_bf16(Nv.ptr(x), B, C, D, CP, Nv.ptr(xt), Nv.stream_ptr()))
    return xt


def from_rows(yt, B, H, D):
    return yt.view(B, D, -1)[:, :, :H].permute(0, 2, 1).contiguous()


def tc_layer(x0t, xkt, W, b, B, m, Hp, H, D, act, direct_begin, fm, col_off, pool=True):
    """One CIN layer through the C ABI (bf16 tensor-core path); returns (y [B,H,D] bf16, yt rows, pooled, maps)."""
    from deepctr import _native as Nv
    L = Nv.lib()
    n = L.xdfm_cin_tc_wprime_elems(m, Hp, H, D)
    assert n > 0, L.xdfm_last_error()
    Hs = (H + 7) // 8 * 8
    wprime = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    yt = torch.full((B * D, Hs), float("nan"), dtype=torch.bfloat16, device=DEV)
    pooled = torch.full((B, fm), float("nan"), device=DEV)
    maps = torch.full((B, fm, D), float("nan"), device=DEV) if not pool else None
    Nv.check(L.xdfm_cin_fwd_tc(Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(W), Nv.ptr(b), Nv.ptr(wprime), B, m, Hp, H, D,
                               Nv.ACT[act], Nv.ptr(yt), direct_begin, Nv.ptr(pooled) if pool else None, Nv.ptr(maps), fm, col_off,
                               Nv.stream_ptr()))
    torch.cuda.synchronize()
    return from_rows(yt, B, H, D), yt, pooled, maps


def emulated_layer(x0b, xkb, W, b, act):
    """fp64 accumulation over bf16-rounded operands and bf16-rounded products (what the kernel computes).
    x0b [B,m,D], xkb [B,Hp,D] bf16 (cpu)."""
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
    (700, 10, 8, 36, 10),         # H not a multiple of 8
]


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("case", TC_CASES, ids=[str(c) for c in TC_CASES])
def test_cin_tc_single_layer_matches_emulation(case, cluster):
    from deepctr import _native as Nv
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case))
    x0 = (torch.randn(B, m, D, generator=g) * 0.5)
    Hprev = Hp if Hp == m else 2 * Hp               # xk is a channel-slice of a wider previous layer when Hp != m
    xk_full = x0 if Hp == m else (torch.randn(B, Hprev, D, generator=g) * 0.5)
    K = Hp * m
    W = torch.randn(H, K, generator=g) / K ** 0.5
    b = torch.randn(H, generator=g) * 0.1
    hdb = H // 2
    fm = H - hdb + 3
    x0t = to_rows(x0.to(DEV), (m + 7) // 8 * 8)
    xkt = x0t if Hp == m else to_rows(xk_full.to(DEV), (Hprev + 7) // 8 * 8)
    Wd, bd = W.to(DEV), b.to(DEV)
    Nv.lib().xdfm_cin_tc_set_cluster(cluster)
    try:
        y, yt, pooled, _ = tc_layer(x0t, xkt, Wd, bd, B, m, Hp, H, D, "relu", hdb, fm, 3)
    finally:
        Nv.lib().xdfm_cin_tc_set_cluster(2)
    ref = emulated_layer(x0.to(torch.bfloat16), xk_full[:, :Hp].to(torch.bfloat16), W, b, "relu")
    scale = ref.abs().max().item()
    # bf16 storage of y: 2^-9 relative; accumulation order differences are ~1e-6
    assert_close(y.float(), ref, 6e-3, 1e-3 * scale, "y (bf16)")
    assert_close(pooled[:, 3:], ref[:, hdb:].sum(-1), 2e-4, 2e-4 * scale * D ** 0.5, "pooled (fp32 from accumulators)")
    Hs = (H + 7) // 8 * 8
    if Hs > H:
        assert float(yt[:, H:].float().abs().max()) == 0.0, "padding channels must be zero"


@pytest.mark.parametrize("case", [(2500, 26, 16, 200, 26), (300, 26, 16, 200, 100), (1300, 12, 32, 64, 12)], ids=str)
def test_cin_tc_single_tile_kernel_still_matches(case):
    """Shapes that default to the tile-pair kernel, forced through the single-tile kernel (xdfm_cin_tc_set_pair(0))."""
    from deepctr import _native as Nv
    Nv.lib().xdfm_cin_tc_set_pair(0)
    try:
        test_cin_tc_single_layer_matches_emulation(case, 2)
    finally:
        Nv.lib().xdfm_cin_tc_set_pair(1)


def test_cin_tc_maps_output():
    B, m, D, H, Hp = 21, 6, 16, 32, 6
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    W = torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5
    b = torch.randn(H, generator=g) * 0.1
    x0t = to_rows(x0.to(DEV), 8)
    Wd, bd = W.to(DEV), b.to(DEV)
    y, _, _, maps = tc_layer(x0t, x0t, Wd, bd, B, m, Hp, H, D, "relu", 16, 20, 4, pool=False)
    ref = emulated_layer(x0.to(torch.bfloat16), x0.to(torch.bfloat16), W, b, "relu")
    assert_close(maps[:, 4:20], ref[:, 16:], 2e-4, 2e-4 * ref.abs().max().item(), "maps (fp32)")


def test_cin_tc_rejects_unsupported_dim():
    from deepctr import _native as Nv
    assert Nv.lib().xdfm_cin_tc_wprime_elems(26, 26, 200, 10) < 0
    assert b"unsupported" in Nv.lib().xdfm_last_error()


DX_CASES = [
    # B, m, D, H, Hp
    (8, 2, 16, 16, 2),
    (19, 26, 16, 200, 100),
    (300, 26, 16, 200, 26),
    (33, 26, 8, 128, 128),
    (20, 22, 32, 256, 128),
    (9, 26, 64, 256, 26),
    (2500, 12, 16, 40, 20),
    (2500, 26, 16, 200, 100),     # 313 tiles on 148 CTAs: 2-3 tiles per CTA, the next tile's dY is staged while the current one drains
    (1250, 22, 32, 256, 128),     # same with the widest A tile (16 granules per row warp, 22 fields)
    (5000, 5, 8, 64, 32),         # fewer fields than granules: several tiles per CTA, each staged at its start
    (2500, 26, 16, 200, 26),      # narrow layer, several tiles per CTA: three fields per MMA group, the last group holds two
    (1300, 39, 16, 128, 64),      # two fields per group, odd field count
    (3000, 3, 16, 32, 3),         # fewer fields than a group could hold
    (2100, 7, 16, 64, 7),         # one group holds every field
]


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("case", DX_CASES, ids=[str(c) for c in DX_CASES])
def test_cin_tc_backward_dx_matches_emulation(case, cluster):
    from deepctr import _native as Nv
    L = Nv.lib()
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    r8 = lambda v: (v + 7) // 8 * 8
    r16 = lambda v: (v + 15) // 16 * 16
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    Hprev = Hp if Hp == m else 2 * Hp
    xk_full = x0 if Hp == m else torch.randn(B, Hprev, D, generator=g) * 0.5
    W = torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5
    dy = torch.randn(B, H, D, generator=g)
    x0t = to_rows(x0.to(DEV), r8(m))
    xkt = x0t if Hp == m else to_rows(xk_full.to(DEV), r8(Hprev))
    dyt = to_rows(dy.to(DEV), r8(H))
    Wd = W.to(DEV)
    wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    HpQ, mP = r16(Hp), r8(m)
    dxk = torch.full((B * D, HpQ), float("nan"), device=DEV)
    dx0 = torch.full((2, B * D, mP), float("nan"), device=DEV)      # two planes (one per channel half), overwritten
    L.xdfm_cin_tc_set_cluster(cluster)
    try:
        Nv.check(L.xdfm_cin_bwd_dx_tc(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), xkt.shape[1], Nv.ptr(Wd), Nv.ptr(wt), B, m, Hp, H, D,
                                      Nv.ptr(dxk), Nv.ptr(dx0), Nv.stream_ptr()))
        torch.cuda.synchronize()
    finally:
        L.xdfm_cin_tc_set_cluster(2)
    bf = lambda t: t.to(torch.bfloat16).double()
    dz = torch.einsum("bhd,hij->bijd", bf(dy), bf(W).view(H, Hp, m))            # [B, Hp, m, D]
    ref_dxk = torch.einsum("bijd,bjd->bid", dz, bf(x0))
    ref_dx0 = torch.einsum("bijd,bid->bjd", dz, bf(xk_full[:, :Hp]))
    got_dxk = dxk.view(B, D, HpQ)[:, :, :Hp].permute(0, 2, 1)
    got_dx0 = dx0.sum(0).cpu().view(B, D, mP)[:, :, :m].permute(0, 2, 1)
    # the planes and layer-0 rows through xdfm_cin_dx0_finish: the reference layout [B, m, D]
    fin = torch.full((B, m, D), float("nan"), device=DEV)
    extra = torch.randn(B * D, HpQ, generator=g).to(DEV)
    if HpQ >= mP:
        Nv.check(L.xdfm_cin_dx0_finish(Nv.ptr(dx0), 2, Nv.ptr(extra), HpQ, B, m, D, mP, Nv.ptr(fin), Nv.stream_ptr()))
        want = dx0.sum(0).view(B, D, mP)[:, :, :m].permute(0, 2, 1) + extra.view(B, D, HpQ)[:, :, :m].permute(0, 2, 1)
        assert_close(fin, want, 1e-6, 1e-6, "dx0 finish")
    assert_close(got_dxk, ref_dxk, 1e-3, 1e-3 * ref_dxk.abs().max().item(), "dxk")
    assert_close(got_dx0, ref_dx0, 1e-3, 1e-3 * ref_dx0.abs().max().item(), "dx0")


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("case", [(300, 26, 16, 200, 100), (2500, 12, 16, 40, 20), (33, 22, 32, 256, 128)], ids=str)
def test_cin_tc_backward_dx_writes_the_dy_rows_of_the_layer_below(case, act):
    """xdfm_cin_bwd_dx_tc_dy: instead of fp32 dXk the kernel writes act'(X^{k-1}) * dXk as bf16 into the hidden-half channels of the
    dY rows of the layer below (what cin_dy_rows_cols would have computed from dXk), and xdfm_cin_dy_rows_cols (dnext = NULL,
    pitch -1) completes the direct-connect channels around it."""
    from deepctr import _native as Nv
    L = Nv.lib()
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case) + 7)
    r8 = lambda v: (v + 7) // 8 * 8
    r16 = lambda v: (v + 15) // 16 * 16
    Hprev = 2 * Hp                                      # the layer below: split_half, hidden half = its first Hp channels
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    yprev = torch.randn(B, Hprev, D, generator=g) * 0.5
    W = torch.randn(H, Hp * m, generator=g) / (Hp * m) ** 0.5
    dy = torch.randn(B, H, D, generator=g)
    x0t, xkt, dyt = to_rows(x0.to(DEV), r8(m)), to_rows(yprev.to(DEV), r8(Hprev)), to_rows(dy.to(DEV), r8(H))
    Wd = W.to(DEV)
    wt = torch.empty(L.xdfm_cin_bwd_dx_tc_wt_elems(m, Hp, H, D), dtype=torch.bfloat16, device=DEV)
    HpQ, mP, Hsp = r16(Hp), r8(m), r8(Hprev)
    R = B * D
    dxk = torch.empty((R, HpQ), device=DEV)
    dx0a = torch.empty((2, R, mP), device=DEV)
    dx0b = torch.empty((2, R, mP), device=DEV)
    dy_prev = torch.full((R, Hsp), 7.0, dtype=torch.bfloat16, device=DEV)
    Nv.check(L.xdfm_cin_bwd_dx_tc(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), Hsp, Nv.ptr(Wd), Nv.ptr(wt), B, m, Hp, H, D, Nv.ptr(dxk),
                                  Nv.ptr(dx0a), Nv.stream_ptr()))
    Nv.check(L.xdfm_cin_bwd_dx_tc_dy(Nv.ptr(dyt), Nv.ptr(x0t), Nv.ptr(xkt), Hsp, Nv.ptr(Wd), Nv.ptr(wt), B, m, Hp, H, D, None,
                                     Nv.ptr(dx0b), Nv.ptr(dy_prev), Hsp, act, Nv.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(dx0a, dx0b)
    want = dxk[:, :Hp]
    if act == 1:
        want = torch.where(xkt[:, :Hp].float() > 0, want, torch.zeros_like(want))
    assert torch.equal(dy_prev[:, :Hp], want.to(torch.bfloat16))
    assert bool((dy_prev[:, HpQ:] == 7.0).all())                          # nothing past HpQ is touched
    # the dY kernel of the layer below, two ways: from fp32 dXk, and around the rows the dX kernel wrote
    Hpad = r16(Hprev)
    fm, col_off, db = Hprev - Hp + 5, 3, Hp
    dpooled = torch.randn(B, fm, generator=g).to(DEV)
    out = []
    for ready in (False, True):
        d_rows = dy_prev.clone() if ready else torch.empty((R, Hsp), dtype=torch.bfloat16, device=DEV)
        d_cols = torch.empty((Hpad, R), dtype=torch.bfloat16, device=DEV)
        Nv.check(L.xdfm_cin_dy_rows_cols(Nv.ptr(xkt), B, D, Hprev, Hsp, Hpad, db, Nv.ptr(dpooled), None, fm, col_off,
                                         None if ready else Nv.ptr(dxk), -1 if ready else HpQ, Hp, act, Nv.ptr(d_rows), Nv.ptr(d_cols),
                                         Nv.stream_ptr()))
        out.append((d_rows, d_cols))
    torch.cuda.synchronize()
    assert torch.equal(out[0][0][:, :Hprev], out[1][0][:, :Hprev])
    assert torch.equal(out[0][1][:Hprev], out[1][1][:Hprev])
    # the same kernel also returns the bias gradient: column sums of the bf16 dY it wrote
    d_rows = dy_prev.clone()
    d_cols = torch.empty((Hpad, R), dtype=torch.bfloat16, device=DEV)
    dbias = torch.full((Hprev,), float("nan"), device=DEV)
    ws = torch.empty(L.xdfm_cin_dy_db_workspace_bytes(B, D, Hpad), dtype=torch.uint8, device=DEV)
    Nv.check(L.xdfm_cin_dy_rows_cols_db(Nv.ptr(xkt), B, D, Hprev, Hsp, Hpad, db, Nv.ptr(dpooled), None, fm, col_off, None, -1, Hp, act,
                                        Nv.ptr(d_rows), Nv.ptr(d_cols), Nv.ptr(dbias), Nv.ptr(ws), ws.numel(), Nv.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(d_cols[:Hprev], out[1][1][:Hprev])
    want_db = d_cols[:Hprev].double().sum(1)
    assert_close(dbias, want_db, 1e-5, 1e-5 * float(d_cols[:Hprev].double().abs().sum(1).max()), "db")


def test_cin_dy_rows():
    from deepctr import _native as Nv
    L = Nv.lib()
    B, D, H, hdb, fm, col_off, n_next = 7, 16, 20, 10, 17, 3, 10
    g = torch.Generator().manual_seed(5)
    y = torch.relu(torch.randn(B, H, D, generator=g))
    dpool = torch.randn(B, fm, generator=g)
    dnext = torch.randn(B, n_next, D, generator=g)
    yt = to_rows(y.to(DEV), 24)
    dnext_rows = dnext.permute(0, 2, 1).reshape(B * D, n_next).contiguous()
    dnext_pad = torch.zeros(B * D, 16)
    dnext_pad[:, :n_next] = dnext_rows
    dyt = torch.full((B * D, 24), float("nan"), dtype=torch.bfloat16, device=DEV)
    dpd, dnd = dpool.to(DEV), dnext_pad.to(DEV)
    Nv.check(L.xdfm_cin_dy_rows(Nv.ptr(yt), B, D, H, 24, hdb, Nv.ptr(dpd), None, fm, col_off, Nv.ptr(dnd), 16, n_next, 1, Nv.ptr(dyt),
                                Nv.stream_ptr()))
    ref = torch.zeros(B, H, D)
    ref[:, hdb:] += dpool[:, col_off:col_off + H - hdb, None]
    ref[:, :n_next] += dnext
    ref = ref * (y.to(torch.bfloat16).float() > 0)
    got = from_rows(dyt, B, H, D).float().cpu()
    assert_close(got, ref.to(torch.bfloat16).float(), 0, 0, "dy rows")
    assert float(dyt[:, H:].float().abs().max()) == 0.0


@pytest.mark.parametrize("B,D,H,n_next,use_maps", [(7, 16, 20, 10, False), (300, 16, 200, 100, False), (33, 8, 36, 0, True), (5, 32, 128, 64, False)])
def test_cin_dy_rows_cols_fused_equals_two_kernels(B, D, H, n_next, use_maps):
    """The fused kernel must write exactly what cin_dy_rows followed by rows_to_cols_bf16 write (both layouts, padding zero)."""
    from deepctr import _native as Nv
    L = Nv.lib()
    hdb, col_off = H - H // 2, 3
    fm = col_off + (H - hdb) + 2
    Hs, H_pad, R = (H + 7) // 8 * 8, (H + 15) // 16 * 16, B * D
    g = torch.Generator().manual_seed(B + H)
    yt = to_rows(torch.relu(torch.randn(B, H, D, generator=g)).to(DEV), Hs)
    dpool = torch.randn(B, fm, generator=g).to(DEV)
    dmaps = torch.randn(B, fm, D, generator=g).to(DEV)
    npitch = max((n_next + 15) // 16 * 16, 16)
    dnext = torch.randn(R, npitch, generator=g).to(DEV)
    args = (Nv.ptr(None if use_maps else dpool), Nv.ptr(dmaps if use_maps else None), fm, col_off, Nv.ptr(dnext if n_next else None),
            npitch, n_next, 1)
    dyt_a = torch.full((R, Hs), float("nan"), dtype=torch.bfloat16, device=DEV)
    Nv.check(L.xdfm_cin_dy_rows(Nv.ptr(yt), B, D, H, Hs, hdb, *args, Nv.ptr(dyt_a), Nv.stream_ptr()))
    dyT_a = to_cols(dyt_a, H, H_pad)
    dyt_b = torch.full((R, Hs), float("nan"), dtype=torch.bfloat16, device=DEV)
    dyT_b = torch.full((H_pad, R), float("nan"), dtype=torch.bfloat16, device=DEV)
    Nv.check(L.xdfm_cin_dy_rows_cols(Nv.ptr(yt), B, D, H, Hs, H_pad, hdb, *args, Nv.ptr(dyt_b), Nv.ptr(dyT_b), Nv.stream_ptr()))
    assert torch.equal(dyt_a.view(torch.int16), dyt_b.view(torch.int16))
    assert torch.equal(dyT_a.view(torch.int16), dyT_b.view(torch.int16))


def to_cols(rows, C, CP):
    from deepctr import _native as Nv
    R = rows.shape[0]
    out = torch.full((CP, R), float("nan"), dtype=torch.bfloat16, device=DEV)
    Nv.check(Nv.lib().xdfm_rows_to_cols_bf16(Nv.ptr(rows), rows.shape[1], R, C, CP, Nv.ptr(out), Nv.stream_ptr()))
    return out


DW_CASES = [
    # B, m, D, H, Hp
    (8, 2, 16, 16, 2),
    (19, 26, 16, 200, 100),
    (300, 26, 16, 200, 26),
    (33, 26, 8, 128, 128),
    (20, 22, 32, 256, 128),
    (9, 25, 64, 256, 26),
    (2500, 12, 16, 40, 20),
    # lane packing (HpQ <= 64): four fields per accumulator (HpQ <= 32) / two (HpQ <= 64); field counts that leave partial groups
    (700, 26, 16, 200, 26),       # cfg2 layer 0: 26 fields = 3 CTAs x 8 + one with 2
    (300, 22, 32, 256, 22),       # cfg4 layer 0: one accumulator per CTA (H_pad = 256), 4 fields each, 22 = 5 x 4 + 2
    (257, 9, 16, 48, 9),          # HpQ = 16 < lane width 32
    (300, 13, 8, 64, 40),         # HpQ = 48: two fields per accumulator, 13 fields
    (120, 5, 16, 24, 64),         # HpQ = 64 exactly
]


@pytest.mark.parametrize("case", [c for c in DW_CASES if (c[4] + 15) // 16 * 16 <= 64], ids=str)
def test_cin_tc_backward_dw_lane_packing_equals_one_field_per_accumulator(case):
    """Packing only changes which TMEM lanes / CTAs hold a field: with the same r-splits the partial sums are the same numbers."""
    from deepctr import _native as Nv
    L = Nv.lib()
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case) + 5)
    r8 = lambda v: (v + 7) // 8 * 8
    r16 = lambda v: (v + 15) // 16 * 16
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    xk_full = x0 if Hp == m else torch.randn(B, 2 * Hp, D, generator=g) * 0.5
    dy = torch.randn(B, H, D, generator=g)
    x0t = to_rows(x0.to(DEV), r8(m))
    xkt = x0t if Hp == m else to_rows(xk_full.to(DEV), r8(2 * Hp))
    dyt = to_rows(dy.to(DEV), r8(H))
    x0T, xkT, dyT = to_cols(x0t, m, r8(m)), to_cols(xkt, Hp, r16(Hp)), to_cols(dyt, H, r16(H))
    outs = []
    try:
        for pack in (1, 0):
            L.xdfm_cin_dw_set_pack(pack)
            nb = L.xdfm_cin_bwd_dw_tc_workspace_bytes(B, m, Hp, H, D)
            ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
            dW = torch.full((H, Hp * m), float("nan"), device=DEV)
            db = torch.full((H,), float("nan"), device=DEV)
            Nv.check(L.xdfm_cin_bwd_dw_tc(Nv.ptr(dyT), Nv.ptr(xkT), Nv.ptr(x0T), B, m, Hp, H, D, Nv.ptr(dW), Nv.ptr(db), Nv.ptr(ws), nb,
                                          Nv.stream_ptr()))
            torch.cuda.synchronize()
            outs.append((dW, db))
    finally:
        L.xdfm_cin_dw_set_pack(1)
    assert torch.isfinite(outs[0][0]).all()
    # the number of r-splits differs between the geometries (more CTAs per r range without packing): equal up to fp32 re-association
    assert_close(outs[0][0], outs[1][0], 1e-5, 1e-5 * outs[1][0].abs().max().item(), "dW packed vs unpacked")
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("case", DW_CASES, ids=[str(c) for c in DW_CASES])
def test_cin_tc_backward_dw_matches_emulation(case, cluster):
    from deepctr import _native as Nv
    L = Nv.lib()
    B, m, D, H, Hp = case
    g = torch.Generator().manual_seed(sum(case) + 2)
    r8 = lambda v: (v + 7) // 8 * 8
    r16 = lambda v: (v + 15) // 16 * 16
    x0 = torch.randn(B, m, D, generator=g) * 0.5
    Hprev = Hp if Hp == m else 2 * Hp
    xk_full = x0 if Hp == m else torch.randn(B, Hprev, D, generator=g) * 0.5
    dy = torch.randn(B, H, D, generator=g)
    x0t = to_rows(x0.to(DEV), r8(m))
    xkt = x0t if Hp == m else to_rows(xk_full.to(DEV), r8(Hprev))
    dyt = to_rows(dy.to(DEV), r8(H))
    x0T, xkT, dyT = to_cols(x0t, m, r8(m)), to_cols(xkt, Hp, r16(Hp)), to_cols(dyt, H, r16(H))
    assert float(dyT[H:].float().abs().sum()) == 0.0 and float(xkT[Hp:].float().abs().sum()) == 0.0
    dW = torch.full((H, Hp * m), float("nan"), device=DEV)
    db = torch.full((H,), float("nan"), device=DEV)
    L.xdfm_cin_tc_set_cluster(cluster)
    try:
        nb = L.xdfm_cin_bwd_dw_tc_workspace_bytes(B, m, Hp, H, D)
        assert nb > 0, L.xdfm_last_error()
        ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
        Nv.check(L.xdfm_cin_bwd_dw_tc(Nv.ptr(dyT), Nv.ptr(xkT), Nv.ptr(x0T), B, m, Hp, H, D, Nv.ptr(dW), Nv.ptr(db), Nv.ptr(ws), nb,
                                      Nv.stream_ptr()))
        torch.cuda.synchronize()
    finally:
        L.xdfm_cin_tc_set_cluster(2)
    bf = lambda t: t.to(torch.bfloat16).double()
    z = (bf(xk_full[:, :Hp])[:, :, None, :].float() * bf(x0)[:, None, :, :].float()).to(torch.bfloat16).double()    # [B,Hp,m,D]
    ref_dW = torch.einsum("bhd,bijd->hij", bf(dy), z).reshape(H, Hp * m)
    ref_db = bf(dy).sum(dim=(0, 2))
    assert_close(dW, ref_dW, 1e-3, 1e-3 * ref_dW.abs().max().item(), "dW")
    assert_close(db, ref_db, 1e-4, 1e-4 * ref_db.abs().max().item(), "db")
