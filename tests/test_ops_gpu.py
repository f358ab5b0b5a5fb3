"""GPU: every C-ABI kernel family against a plain PyTorch fp32 reference of the same op
(CUDA, TF32 disabled), called through ctypes exactly like the product path does."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _fp32_reference_math():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _lib():
    from dpc_b200._lib import lib
    return lib()


def _st():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.double(), b.double()
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def to_rows(x):      # NCDHW -> [rows, C]
    return x.permute(0, 2, 3, 4, 1).contiguous().view(-1, x.shape[1])


def from_rows(r, NB, T, H, W):
    return r.view(NB, T, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize('NB,T,H,W', [(2, 5, 64, 64), (3, 2, 32, 48), (1, 1, 28, 20), (1, 2, 27, 21)])
def test_stem_conv(NB, T, H, W):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn(NB, 3, T, H, W, device='cuda', generator=g)
    w = torch.randn(64, 3, 1, 7, 7, device='cuda', generator=g) * 0.1
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    xr, wr = x.clone(), w.clone().requires_grad_(True)
    yref = F.conv3d(xr, wr, None, (1, 2, 2), (0, 3, 3))
    assert tuple(yref.shape[2:]) == (T, Ho, Wo)
    # conv1 from the fp32 video (odd-frame fallback) + fused bn1 statistics
    y2 = torch.full((NB * T * Ho * Wo, 64), float('nan'), device='cuda')
    ws = torch.empty(128, dtype=torch.float64, device='cuda')
    L.stem_conv_fwd_tc(x.data_ptr(), w.data_ptr(), y2.data_ptr(), ws.data_ptr(), NB, T, H, W, _st())
    torch.cuda.synchronize()
    assert not torch.isnan(y2).any()
    assert rel(from_rows(y2, NB, T, Ho, Wo), yref) < 5e-5
    mean, rstd = torch.empty(64, device='cuda'), torch.empty(64, device='cuda')
    L.bn_finalize(ws.data_ptr(), y2.shape[0], 64, 1e-5, mean.data_ptr(), rstd.data_ptr(), _st())
    assert float((mean - y2.double().mean(0)).abs().max()) < 1e-5 * float(y2.abs().max())
    assert rel(rstd, 1 / torch.sqrt(y2.double().var(0, unbiased=False) + 1e-5)) < 1e-5
    if H % 2 == 0 and W % 2 == 0:
        # space-to-depth planes (the conv over them is covered by test_stem_pool)
        bf = dict(dtype=torch.bfloat16, device='cuda')
        x2h, x2l = torch.empty(NB, T, H // 2, W // 2, 16, **bf), torch.empty(NB, T, H // 2, W // 2, 16, **bf)
        L.stem_s2d_pack(x.data_ptr(), x2h.data_ptr(), x2l.data_ptr(), NB, T, H, W, _st())
        xs = x.view(NB, 3, T, H // 2, 2, W // 2, 2).permute(0, 2, 3, 5, 1, 4, 6).reshape(NB, T, H // 2, W // 2, 12)
        assert torch.equal(x2h[..., :12], xs.to(torch.bfloat16)) and not x2h[..., 12:].any()
        assert float(((x2h.float() + x2l.float())[..., :12] - xs).abs().max()) < 2.0 ** -15 * float(xs.abs().max())
    dy = torch.randn(yref.shape, device='cuda', generator=g)
    yref.backward(dy)
    from dpc_b200 import engine as E
    dyp = E._split(to_rows(dy), _st())
    dw2 = torch.full_like(w, float('nan'))
    L.stem_conv_wgrad_tc(x.data_ptr(), dyp[0].data_ptr(), dyp[1].data_ptr(), dw2.data_ptr(), NB, T, H, W, _st())
    torch.cuda.synchronize()
    assert not torch.isnan(dw2).any()
    assert rel(dw2, wr.grad) < 5e-5
    if H % 2 == 0 and W % 2 == 0:
        dw3 = torch.full_like(w, float('nan'))
        L.stem_conv_wgrad_s2d(x2h.data_ptr(), x2l.data_ptr(), dyp[0].data_ptr(), dyp[1].data_ptr(), dw3.data_ptr(),
                              NB, T, H, W, _st())
        torch.cuda.synchronize()
        assert not torch.isnan(dw3).any()
        assert rel(dw3, wr.grad) < 5e-5


@pytest.mark.parametrize('C,rows', [(64, 5000), (128, 1237), (256, 96)])
@pytest.mark.parametrize('mode', ['plain', 'res', 'resbn'])
@pytest.mark.parametrize('relu', [True, False])
def test_bn_fwd_bwd(C, rows, mode, relu):
    from dpc_b200 import engine as E
    g = torch.Generator(device='cuda').manual_seed(3)
    y = torch.randn(rows, C, device='cuda', generator=g) * 2 + 0.5
    gamma = torch.rand(C, device='cuda', generator=g) + 0.5
    beta = torch.randn(C, device='cuda', generator=g) * 0.1
    res = torch.randn(rows, C, device='cuda', generator=g)
    gr = torch.rand(C, device='cuda', generator=g) + 0.5
    br = torch.randn(C, device='cuda', generator=g) * 0.1
    st = _st()
    mean, rstd = E._bn_stats(y, rows, C, st)
    assert rel(mean, y.mean(0)) < 1e-5
    assert rel(rstd, 1 / torch.sqrt(y.var(0, unbiased=False) + 1e-5)) < 1e-5
    yr = y.clone().requires_grad_(True)
    gam, bet = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True)
    grr, brr = gr.clone().requires_grad_(True), br.clone().requires_grad_(True)
    ref = F.batch_norm(yr, None, None, gam, bet, True, 0.0, 1e-5)
    if mode == 'plain':
        out, _ = E._bn_apply(y, mean, rstd, gamma, beta, relu, rows, C, st)
    elif mode == 'res':
        out, _ = E._bn_apply(y, mean, rstd, gamma, beta, relu, rows, C, st, res=res)
        ref = ref + rr
    else:
        mr, sr = E._bn_stats(res, rows, C, st)
        out, _ = E._bn_apply(y, mean, rstd, gamma, beta, relu, rows, C, st, res=res, rbn=(mr, sr, gr, br))
        ref = ref + F.batch_norm(rr, None, None, grr, brr, True, 0.0, 1e-5)
    if relu:
        ref = F.relu(ref)
    assert rel(out, ref) < 1e-5
    dout = torch.randn(rows, C, device='cuda', generator=g)
    ref.backward(dout)
    dy, _, dg, db, gbuf = E._bn_bwd(dout, out, relu, y, mean, rstd, gamma, rows, C, st, want_g=True)
    assert rel(dy, yr.grad) < 5e-5
    assert rel(dg, gam.grad) < 5e-5
    assert rel(db, bet.grad) < 5e-5
    if mode == 'res':
        assert rel(gbuf, rr.grad) < 1e-6
    if mode == 'resbn':
        dyr, _, dgr, dbr, _ = E._bn_bwd(dout, out, relu, res, mr, sr, gr, rows, C, st)
        assert rel(dyr, rr.grad) < 5e-5
        assert rel(dgr, grr.grad) < 5e-5


def test_bn_plane_io_matches_row_io():
    """the fused split-bf16 plane outputs / plane residual / hi-plane ReLU mask agree with the fp32-row path"""
    from dpc_b200 import engine as E
    rows, C = 3000, 128
    g = torch.Generator(device='cuda').manual_seed(33)
    y = torch.randn(rows, C, device='cuda', generator=g) * 2 + 0.3
    gamma = torch.rand(C, device='cuda', generator=g) + 0.5
    beta = torch.randn(C, device='cuda', generator=g) * 0.1
    res = torch.randn(rows, C, device='cuda', generator=g)
    st = _st()
    mean, rstd = E._bn_stats(y, rows, C, st)
    ref, _ = E._bn_apply(y, mean, rstd, gamma, beta, True, rows, C, st, res=res)
    resp = E._split(res, st)
    out, pl = E._bn_apply(y, mean, rstd, gamma, beta, True, rows, C, st, res_planes=resp, want_rows=True, want_planes=True)
    assert rel(out, ref) < 2e-5                                   # residual read through planes: ~2^-17
    rec = pl[0].float() + pl[1].float()
    assert rel(rec, out) < 2e-5
    assert torch.equal(pl[0], out.to(torch.bfloat16))
    hi2, lo2 = E._split(out, st)
    assert torch.equal(hi2, pl[0]) and torch.equal(lo2, pl[1])    # same rounding as the stand-alone split kernel
    dout = torch.randn(rows, C, device='cuda', generator=g)
    dy, _, dg, db, gb = E._bn_bwd(dout, out, True, y, mean, rstd, gamma, rows, C, st, want_g=True)
    _, dyp, dg2, db2, gb2 = E._bn_bwd(dout, None, True, y, mean, rstd, gamma, rows, C, st, want_g=True, out_hi=pl[0],
                                      want_rows=False, want_planes=True)
    assert torch.equal(gb, gb2)                                   # identical ReLU mask from the hi plane
    assert rel(dg2, dg) < 1e-6 and rel(db2, db) < 1e-6            # (fp64 atomics: summation order varies)
    assert rel(dyp[0].float() + dyp[1].float(), dy) < 2e-5


@pytest.mark.parametrize('NT,H,W', [(6, 32, 32), (3, 16, 24), (2, 7, 9)])
def test_bn_relu_maxpool(NT, H, W):
    from dpc_b200 import engine as E
    L = _lib()
    C = 64
    g = torch.Generator(device='cuda').manual_seed(4)
    y = torch.randn(NT * H * W, C, device='cuda', generator=g)
    gamma = torch.rand(C, device='cuda', generator=g) + 0.5
    beta = torch.randn(C, device='cuda', generator=g) * 0.1
    st = _st()
    mean, rstd = E._bn_stats(y, NT * H * W, C, st)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(NT * Ho * Wo, C, device='cuda')
    L.bn_relu_maxpool_fwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                          out.data_ptr(), NT, H, W, C, st)
    # reference: treat NT as the batch, one "frame" each
    yr = y.view(NT, H, W, C).permute(0, 3, 1, 2).unsqueeze(2).contiguous().requires_grad_(True)   # [NT,C,1,H,W]
    gam, bet = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a = F.relu(F.batch_norm(yr, None, None, gam, bet, True, 0.0, 1e-5))
    a.retain_grad()
    ref = F.max_pool3d(a, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert tuple(ref.shape[3:]) == (Ho, Wo)
    ref_rows = ref.squeeze(2).permute(0, 2, 3, 1).reshape(-1, C)
    assert rel(out, ref_rows) < 1e-5
    dout = torch.randn(NT * Ho * Wo, C, device='cuda', generator=g)
    ref.backward(dout.view(NT, Ho, Wo, C).permute(0, 3, 1, 2).unsqueeze(2))
    gbuf = torch.empty_like(y)
    L.bn_relu_maxpool_bwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                          out.data_ptr(), dout.data_ptr(), gbuf.data_ptr(), NT, H, W, C, st)
    # gbuf = grad wrt bn(y) output (ReLU mask applied); compare through the BN backward
    dy, _, dg, db, _ = E._bn_bwd(gbuf, None, False, y, mean, rstd, gamma, NT * H * W, C, st)
    dy_ref = yr.grad.squeeze(2).permute(0, 2, 3, 1).reshape(-1, C)
    assert rel(dy, dy_ref) < 5e-5
    assert rel(dg, gam.grad) < 5e-5
    # fused stem tail backward (no materialised g)
    ws = torch.empty(2 * C, dtype=torch.float64, device='cuda')
    dg2, db2, dy2 = torch.empty(C, device='cuda'), torch.empty(C, device='cuda'), torch.empty_like(y)
    hi = torch.empty(y.shape, dtype=torch.bfloat16, device='cuda')
    lo = torch.empty_like(hi)
    for pooled in (0, 1):
        L.stem_tail_bwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                        dout.data_ptr(), ws.data_ptr(), dg2.data_ptr(), db2.data_ptr(), dy2.data_ptr(), hi.data_ptr(),
                        lo.data_ptr(), NT, H, W, C, pooled, st)
        assert rel(dy2, dy_ref) < 5e-5 and rel(dg2, gam.grad) < 5e-5 and rel(db2, bet.grad) < 5e-5, pooled
        assert rel(hi.float() + lo.float(), dy2) < 2e-5


def test_pool_split():
    L = _lib()
    NB, T, S, C = 6, 2, 16, 256
    g = torch.Generator(device='cuda').manual_seed(5)
    z = torch.randn(NB, T, S, C, device='cuda', generator=g)
    finf, feat = torch.empty(NB, S, C, device='cuda'), torch.empty(NB, S, C, device='cuda')
    L.pool_split_fwd(z.data_ptr(), finf.data_ptr(), feat.data_ptr(), NB, T, S, C, _st())
    zr = z.clone().requires_grad_(True)
    m = zr.mean(1)
    assert rel(finf, m) < 1e-6 and rel(feat, F.relu(m)) < 1e-6
    d1, d2 = torch.randn_like(finf), torch.randn_like(finf)
    (m * d1 + F.relu(m) * d2).sum().backward()
    dz = torch.empty_like(z)
    L.pool_split_bwd(finf.data_ptr(), d1.data_ptr(), d2.data_ptr(), dz.data_ptr(), NB, T, S, C, _st())
    assert rel(dz, zr.grad) < 1e-6


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(200, 256, 512), (96, 96, 256), (33, 70, 19), (2048, 256, 256), (256, 256, 2048),
                                   (128, 512, 64)])
def test_gemm_f32(ta, tb, M, N, K):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(6)
    A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
    B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
    C0 = torch.randn(M, N, device='cuda', generator=g)
    C = C0.clone()
    L.gemm_f32(ta, tb, M, N, K, 0.5, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], 2.0, C.data_ptr(), N, _st())
    ref = 0.5 * ((A.t() if ta else A).double() @ (B.t() if tb else B).double()) + 2.0 * C0.double()
    assert rel(C, ref) < 1e-5


def test_rows_and_colsum():
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(7)
    B, N, S, D, P = 3, 8, 4, 256, 3
    feat = torch.randn(B * N * S, D, device='cuda', generator=g)
    for t in (0, 4):
        dst = torch.empty(B * S, D, device='cuda')
        L.gather_rows(feat.data_ptr(), dst.data_ptr(), B * S, D, S, N * S, t * S, _st())
        assert torch.equal(dst, feat.view(B, N, S, D)[:, t].reshape(B * S, D))
    dst = torch.empty(B * P * S, D, device='cuda')
    L.gather_rows(feat.data_ptr(), dst.data_ptr(), B * P * S, D, P * S, N * S, (N - P) * S, _st())
    assert torch.equal(dst, feat.view(B, N, S, D)[:, N - P:].reshape(-1, D))
    back = torch.zeros_like(feat)
    L.scatter_rows(dst.data_ptr(), back.data_ptr(), B * P * S, D, P * S, N * S, (N - P) * S, 0, _st())
    assert torch.equal(back.view(B, N, S, D)[:, N - P:].reshape(-1, D), dst)
    assert float(back.view(B, N, S, D)[:, :N - P].abs().max()) == 0.0
    A = torch.randn(1000, 512, device='cuda', generator=g)
    out = torch.ones(512, device='cuda')
    L.colsum(A.data_ptr(), 1000, 512, out.data_ptr(), 1, _st())
    assert rel(out, A.double().sum(0) + 1) < 1e-5


@pytest.mark.parametrize('B,P,SQ', [(2, 3, 4), (3, 2, 9), (5, 3, 16)])
def test_nce_mask_and_ce(B, P, SQ):
    from dpc_b200 import engine as E
    from oracle import dpc_oracle as O
    m = E.nce_mask(B, P, SQ, torch.device('cuda'))
    ref = O.closed_form_mask(B, P, int(math.isqrt(SQ)))
    assert m.dtype == torch.int8 and m.is_contiguous()
    assert torch.equal(m.cpu(), ref)
    M = B * P * SQ
    g = torch.Generator(device='cuda').manual_seed(8)
    for rows in (M, 2 * M):                        # square (1 GPU) and DataParallel-gathered (2 replicas)
        score = (torch.randn(rows, M, device='cuda', generator=g) * 3).requires_grad_(True)
        target = torch.arange(rows, device='cuda') % M
        ref_loss = F.cross_entropy(score, target)
        ref_loss.backward()
        out, lse = E.nce_ce_forward(score.detach())
        assert abs(float(out[0]) - float(ref_loss.detach())) < 1e-5 * max(1.0, abs(float(ref_loss.detach())))
        tk = O.topk_accuracy(score.detach(), target)
        for i in range(3):
            assert abs(float(out[1 + i]) - float(tk[i])) < 1e-6
        gs = torch.tensor([0.7], device='cuda')
        d = E.nce_ce_backward(score.detach(), lse, gs)
        assert rel(d, 0.7 * score.grad) < 1e-5


def test_adam_step():
    from oracle import dpc_oracle as O
    L = _lib()
    n = 100003
    g = torch.Generator(device='cuda').manual_seed(9)
    p = torch.randn(n, device='cuda', generator=g)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pref = {'p': p.clone().cpu()}
    st = {}
    for step in range(1, 4):
        grad = torch.randn(n, device='cuda', generator=g)
        L.adam_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-5,
                    step, 0.5, _st())
        O.adam_step(pref, {'p': (grad * 0.5).cpu()}, st)
    assert rel(p.cpu(), pref['p']) < 1e-6


def test_gru_cell_matches_oracle():
    """one fused ConvGRU step (GEMMs + gate kernels) vs oracle.gru_cell, eval mode"""
    from dpc_b200 import engine as E
    from oracle import dpc_oracle as O
    B, L_, D = 3, 2, 256
    sd = O.synthetic_state_dict('resnet18', 5)
    P = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, D, L_, L_, generator=g)
    h = torch.randn(B, D, L_, L_, generator=g)
    ref = O.gru_cell(x, h, sd)
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, D).contiguous().cuda()
    gru = E._Gru(P, D, _st())
    xr, hr = rows(x), rows(h)
    XP = gru.xproj(xr, B * L_ * L_)
    hn, _ = gru.step(XP, 0, hr, B * L_ * L_, 0.0, 0, 0)
    assert rel(hn.cpu(), rows(ref).cpu()) < 2e-5


@pytest.mark.parametrize('NB,T,H,W', [(2, 5, 64, 64), (2, 2, 128, 128), (1, 2, 64, 224), (3, 1, 36, 20), (1, 1, 90, 44),
                                      (1, 1, 224, 224), (5, 5, 128, 128), (64, 5, 128, 128)])
def test_stem_pool(NB, T, H, W):
    """pooled stem (stem_pool.cu): conv1 + bn1 statistics + selected pooled value / window index, finalize into operand planes,
    pooled-grid BatchNorm-backward sums, recomputing backward -> gradient planes on the conv1 grid.  One band (<= ~80 conv
    columns) and two bands (W = 224), odd pooled extents, several frames per CTA, negative gamma (min-pool channels)."""
    L = _lib()
    assert L.stem_pool_supported(H, W) >= 1
    g = torch.Generator(device='cuda').manual_seed(12)
    x = torch.randn(NB, 3, T, H, W, device='cuda', generator=g)
    w = torch.randn(64, 3, 1, 7, 7, device='cuda', generator=g) * 0.1
    gamma = (torch.rand(64, device='cuda', generator=g) + 0.5) * torch.where(torch.rand(64, device='cuda', generator=g) < 0.25, -1.0, 1.0)
    beta = torch.randn(64, device='cuda', generator=g) * 0.3
    Ho, Wo = H // 2, W // 2
    Hp, Wp = (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1
    rows0, rows_p = NB * T * Ho * Wo, NB * T * Hp * Wp
    bf = dict(dtype=torch.bfloat16, device='cuda')
    x2h, x2l = torch.empty(NB, T, Ho, Wo, 16, **bf), torch.empty(NB, T, Ho, Wo, 16, **bf)
    wp = torch.empty(32768, **bf)
    L.stem_s2d_pack(x.data_ptr(), x2h.data_ptr(), x2l.data_ptr(), NB, T, H, W, _st())
    L.stem_s2d_wpack(w.data_ptr(), wp.data_ptr(), _st())
    ypool = torch.full((rows_p, 64), float('nan'), device='cuda')
    idx = torch.full((rows_p, 64), 255, dtype=torch.uint8, device='cuda')
    ws = torch.empty(128, dtype=torch.float64, device='cuda')
    L.stem_pool_fwd(x2h.data_ptr(), x2l.data_ptr(), wp.data_ptr(), gamma.data_ptr(), ypool.data_ptr(), idx.data_ptr(),
                    ws.data_ptr(), NB, T, H, W, _st())
    torch.cuda.synchronize()
    assert not torch.isnan(ypool).any() and int(idx.max()) <= 8
    y0 = F.conv3d(x, w, None, (1, 2, 2), (0, 3, 3))                                   # [NB,64,T,Ho,Wo]
    y0r = to_rows(y0)                                                                 # [rows0, 64]
    ymax = float(y0.abs().max())
    mean, rstd = torch.empty(64, device='cuda'), torch.empty(64, device='cuda')
    L.bn_finalize(ws.data_ptr(), rows0, 64, 1e-5, mean.data_ptr(), rstd.data_ptr(), _st())
    assert float((mean - y0r.double().mean(0)).abs().max()) < 1e-5 * ymax
    assert rel(rstd, 1 / torch.sqrt(y0r.double().var(0, unbiased=False) + 1e-5)) < 1e-5
    # the kept value is the conv output at the kept window index, and it is the window's max of sign(gamma) * y
    y5 = y0r.view(NB * T, Ho, Wo, 64)
    ypad = torch.full((NB * T, 2 * Hp + 1, 2 * Wp + 1, 64), float('nan'), device='cuda')
    ypad[:, 1:Ho + 1, 1:Wo + 1] = y5
    win = torch.stack([ypad[:, dh:dh + 2 * Hp:2, dw:dw + 2 * Wp:2] for dh in range(3) for dw in range(3)], -1)   # [NT,Hp,Wp,64,9]
    kept = torch.gather(win, -1, idx.view(NB * T, Hp, Wp, 64, 1).long()).squeeze(-1)
    assert not torch.isnan(kept).any()                                                # never a padding position
    assert float((kept.reshape(-1, 64) - ypool).abs().max()) < 2e-5 * ymax
    sgn = torch.where(gamma < 0, -1.0, 1.0)
    wmax = torch.nan_to_num(win * sgn.view(1, 1, 1, 64, 1), nan=-1e30).max(-1).values
    assert float((wmax.reshape(-1, 64) - ypool * sgn).abs().max()) < 2e-5 * ymax
    # finalize: operand planes of relu(bn1(.)) on the pooled grid == maxpool(relu(bn1(conv1)))
    ah, al = torch.empty(rows_p, 64, **bf), torch.empty(rows_p, 64, **bf)
    arows = torch.empty(rows_p, 64, device='cuda')
    L.stem_pool_finalize(ypool.data_ptr(), idx.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                         ah.data_ptr(), al.data_ptr(), arows.data_ptr(), rows_p, _st())
    a_ref = F.max_pool3d(F.relu(F.batch_norm(y0, None, None, gamma, beta, True, 0.0, 1e-5)), (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert tuple(a_ref.shape[2:]) == (T, Hp, Wp)
    a_ref = to_rows(a_ref)
    amax = float(a_ref.abs().max())
    assert float((arows - a_ref).abs().max()) < 5e-5 * amax
    assert float((ah.float() + al.float() - arows).abs().max()) < 2e-5 * amax
    assert torch.equal(ah, arows.to(torch.bfloat16))
    assert torch.equal((idx & 0x80) != 0, arows <= 0)                                 # ReLU-dead windows flagged
    # backward: reference built with OUR window choice (a near-tie may legitimately pick the other position)
    dout = torch.randn(rows_p, 64, device='cuda', generator=g)
    k = idx.view(NB * T, Hp, Wp, 64)
    d5 = dout.view(NB * T, Hp, Wp, 64).double()
    gpad = torch.zeros(NB * T, 2 * Hp + 1, 2 * Wp + 1, 64, dtype=torch.float64, device='cuda')
    for dh in range(3):
        for dw in range(3):
            gpad[:, dh:dh + 2 * Hp:2, dw:dw + 2 * Wp:2] += d5 * (k == dh * 3 + dw)
    gg = gpad[:, 1:Ho + 1, 1:Wo + 1].reshape(rows0, 64)
    assert float(gpad.sum() - gg.sum()) == 0.0 or abs(float(gpad.sum() - gg.sum())) < 1e-9 * float(gg.abs().sum())
    xhat = (y0r.double() - mean.double()) * rstd.double()
    sg, sgx = gg.sum(0), (gg * xhat).sum(0)
    dy_ref = gamma.double() * rstd.double() * (gg - sg / rows0 - xhat * sgx / rows0)
    ws2 = torch.empty(128, dtype=torch.float64, device='cuda')
    dga, dbe = torch.empty(64, device='cuda'), torch.empty(64, device='cuda')
    L.stem_pool_bwd_reduce(ypool.data_ptr(), dout.data_ptr(), idx.data_ptr(), mean.data_ptr(), rstd.data_ptr(), ws2.data_ptr(),
                           dga.data_ptr(), dbe.data_ptr(), rows_p, _st())
    assert float((dbe.double() - sg).abs().max()) < 5e-5 * float(sg.abs().max())
    assert float((dga.double() - sgx).abs().max()) < 5e-5 * float(sgx.abs().max())
    dyh = torch.full((rows0, 64), float('nan'), **bf)
    dyl = torch.full((rows0, 64), float('nan'), **bf)
    L.stem_pool_bwd(x2h.data_ptr(), x2l.data_ptr(), wp.data_ptr(), dout.data_ptr(), idx.data_ptr(), mean.data_ptr(),
                    rstd.data_ptr(), gamma.data_ptr(), ws2.data_ptr(), dyh.data_ptr(), dyl.data_ptr(), NB, T, H, W, _st())
    torch.cuda.synchronize()
    dy = dyh.float() + dyl.float()
    assert not torch.isnan(dy).any()
    assert float((dy.double() - dy_ref).abs().max()) < 5e-5 * float(dy_ref.abs().max())
    # the same backward with conv1's wgrad fused into the kernel (gradient tile in shared memory -> MN-major UMMA operand)
    if L.stem_pool_supported(H, W) == 2:
        wr = w.clone().requires_grad_(True)
        F.conv3d(x, wr, None, (1, 2, 2), (0, 3, 3)).backward(from_rows(dy_ref.float(), NB, T, Ho, Wo))
        dw = torch.full_like(w, float('nan'))
        L.stem_pool_bwd_wgrad(x2h.data_ptr(), x2l.data_ptr(), wp.data_ptr(), dout.data_ptr(), idx.data_ptr(), mean.data_ptr(),
                              rstd.data_ptr(), gamma.data_ptr(), ws2.data_ptr(), dw.data_ptr(), NB, T, H, W, _st())
        torch.cuda.synchronize()
        assert not torch.isnan(dw).any()
        assert rel(dw, wr.grad) < 5e-5


@pytest.mark.parametrize('B,L_,p', [(3, 2, 0.0), (4, 4, 0.1), (5, 7, 0.1)])
def test_head_chain_matches_stepwise(B, L_, p):
    """head_chain.cu (GRU aggregation + prediction loop as one kernel per direction, tensor-core weight gradients) against the
    per-step GEMM + gate kernels (which tests/test_parity_gpu.py pins to the oracle): same dropout stream, so train mode too.
    L_ = 7 is the 224^2 extent (rows per clip 49: the last 16-row CTA is partial)."""
    from dpc_b200 import engine as E
    from oracle import dpc_oracle as O
    N, P_, D, To = 8, 3, 256, 2
    S = L_ * L_
    sd = O.synthetic_state_dict('resnet18', 17)
    P = {k: v.cuda() for k, v in sd.items() if k.startswith('agg.cell_list') or k.startswith('network_pred')}
    g = torch.Generator(device='cuda').manual_seed(18)
    z4 = torch.randn(B * N * To * S, D, device='cuda', generator=g)
    M = B * P_ * S
    dscore = torch.randn(M, M, device='cuda', generator=g) / M
    out = {}
    for chain in (False, True):
        E.HEAD_CHAIN = chain
        try:
            score, ctx = E.head_forward(z4, (To, L_, L_), B, N, P_, P, dropout_p=p, seed=4242)
            assert bool(ctx.get('chain', False)) == chain
            dz4, G = E.head_backward(ctx, dscore.clone(), P)
            torch.cuda.synchronize()
        finally:
            E.HEAD_CHAIN = True
        out[chain] = (score, dz4, G)
    s0, d0, G0 = out[False]
    s1, d1, G1 = out[True]
    assert rel(s1, s0) < 1e-5
    assert rel(d1, d0) < 2e-5
    for k in E.HEAD_PARAM_NAMES:
        assert G1[k].shape == G0[k].shape, k
        assert rel(G1[k], G0[k]) < 1e-4, (k, rel(G1[k], G0[k]))
