"""GPU: the tcgen05 (3xBF16 split) kernels against fp32/fp64 PyTorch references, through the C ABI."""
import math

import pytest
import torch
import torch.nn.functional as F

from dpc_b200._lib import ConvGeom

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
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def split(x, f16=False):
    """bf16 pairs (dpc_split_bf16: every conv operand) or fp16 pairs (dpc_split_f16: forward-value-only GEMMs)"""
    x = x.contiguous()
    dt = torch.float16 if f16 else torch.bfloat16
    hi = torch.empty(x.shape, dtype=dt, device=x.device)
    lo = torch.empty(x.shape, dtype=dt, device=x.device)
    (_lib().split_f16 if f16 else _lib().split_bf16)(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _st())
    return hi, lo


def to_rows(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().view(-1, x.shape[1])


def from_rows(r, NB, T, H, W):
    return r.view(NB, T, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


def test_split_planes():
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(1 << 16, device='cuda', generator=g) * 3
    hi, lo = split(x)                                                 # bf16 pairs: 16 mantissa bits
    rec = hi.float() + lo.float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-20)).max()) < 2.0 ** -15
    assert torch.equal(hi, x.to(torch.bfloat16))
    hi, lo = split(x, f16=True)                                       # fp16 pairs: 22 mantissa bits (values O(1))
    rec = hi.float() + lo.float()
    big = x.abs() > 0.25                       # below that the lo plane is an fp16 subnormal (absolute precision 3e-8)
    assert float(((rec - x).abs() / x.abs())[big].max()) < 2.0 ** -20
    assert float((rec - x).abs().max()) < 2.0 ** -20 * 16
    assert torch.equal(hi, x.to(torch.float16))


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (256, 256, 256), (1024, 768, 256), (200, 100, 128),
                                   (1617, 1617, 256), (6144, 512, 256)])
@pytest.mark.parametrize('f16', [0, 1])
def test_gemm_nt_split(M, N, K, f16):
    """f16 = 1: both operands fp16 pairs (the score matmul's forward: 5e-6); 0: bf16 pairs (its backward GEMMs: 5e-5)"""
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(2)
    A = torch.randn(M, K, device='cuda', generator=g)
    B = torch.randn(N, K, device='cuda', generator=g)
    ah, al = split(A, f16=bool(f16))
    bh, bl = split(B, f16=bool(f16))
    C = torch.full((M, N), float('nan'), device='cuda')
    L.gemm_nt_split_tc(M, N, K, ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), f16, C.data_ptr(), 0, _st())
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    assert not torch.isnan(C).any()
    tol = 5e-6 if f16 else 5e-5
    assert rel(C, ref) < tol
    C2 = torch.ones(M, N, device='cuda')
    L.gemm_nt_split_tc(M, N, K, ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), f16, C2.data_ptr(), 1, _st())
    assert rel(C2, ref + 1) < tol


@pytest.mark.parametrize('M,N', [(128, 256), (6144, 6144), (1617, 1617), (6468, 6468), (200, 100), (36, 36)])
@pytest.mark.parametrize('f16', [1, 0])
def test_score_matmul(M, N, f16):
    """persistent A-resident score matmul (score_tc.cu), K = 256: config-2 / config-4 / config-5 sizes, partial tiles, odd N"""
    L = _lib()
    K = 256
    g = torch.Generator(device='cuda').manual_seed(5)
    A = torch.randn(M, K, device='cuda', generator=g) * 0.25
    B = torch.randn(N, K, device='cuda', generator=g)
    ah, al = split(A, f16=bool(f16))
    bh, bl = split(B, f16=bool(f16))
    C = torch.full((M, N), float('nan'), device='cuda')
    L.score_matmul_tc(M, N, K, ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), f16, C.data_ptr(), _st())
    torch.cuda.synchronize()
    assert not torch.isnan(C).any()
    ref = A.double() @ B.double().t()
    assert rel(C, ref) < (5e-6 if f16 else 5e-5)


CASES = [
    # NB, T, H, W, Ci, Co, k, s, p
    (2, 5, 32, 32, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (3, 5, 16, 16, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (4, 3, 8, 8, 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (8, 2, 4, 4, 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (5, 2, 7, 7, 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (3, 3, 14, 14, 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    (2, 5, 28, 28, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (3, 1, 8, 8, 64, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    # strided sites of the backbone
    (2, 5, 32, 32, 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),       # layer2.0.conv1
    (2, 5, 32, 32, 64, 128, (1, 1, 1), (1, 2, 2), (0, 0, 0)),       # layer2.0.downsample
    (3, 5, 16, 16, 128, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),      # layer3.0.conv1
    (3, 5, 16, 16, 128, 256, (1, 1, 1), (2, 2, 2), (0, 0, 0)),      # layer3.0.downsample
    (4, 3, 8, 8, 256, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),        # layer4.0.conv1
    (3, 3, 7, 7, 256, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),        # odd extents
    (2, 5, 28, 28, 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    # >= 2 tiles per SM with narrow outputs: the persistent kernel (resident / streamed weights)
    (8, 5, 32, 32, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (32, 5, 16, 16, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (32, 5, 32, 32, 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (32, 5, 32, 32, 64, 128, (1, 1, 1), (1, 2, 2), (0, 0, 0)),
    (9, 5, 30, 30, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    # 64 -> 64 frames too wide for the halo-patch kernels' shared-memory budget: tap-per-box fallback
    (1, 2, 56, 56, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    # Bottleneck2d / 3d sites (resnet_2d3d.py:119-202; every Ci / Co / stride combination of backbone_spec('resnet50')):
    # 1x1x1 reductions / expansions up to 1024 channels (grid.y > 1 with BN = 256, wgrad with Ci = 1024), strided
    # 1x1x1 downsamples, the strided 3x3 after a 1x1x1
    (3, 5, 16, 16, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),        # layer1.0.conv1
    (3, 5, 16, 16, 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # layer1.x.conv3 / layer1.0.downsample
    (3, 5, 16, 16, 256, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # layer1.1.conv1
    (3, 5, 16, 16, 256, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0)),      # layer2.0.conv1
    (3, 5, 16, 16, 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),      # layer2.0.conv2 (carries the stride)
    (3, 5, 8, 8, 128, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0)),        # layer2.x.conv3
    (3, 5, 16, 16, 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0)),      # layer2.0.downsample
    (3, 5, 8, 8, 512, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0)),        # layer2.1.conv1
    (3, 5, 8, 8, 512, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0)),        # layer3.0.conv1
    (3, 5, 8, 8, 256, 256, (3, 3, 3), (2, 2, 2), (1, 1, 1)),        # layer3.0.conv2
    (3, 3, 4, 4, 256, 1024, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # layer3.x.conv3
    (3, 5, 8, 8, 512, 1024, (1, 1, 1), (2, 2, 2), (0, 0, 0)),       # layer3.0.downsample
    (3, 3, 4, 4, 1024, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # layer3.1.conv1 / layer4.0.conv1
    (3, 3, 4, 4, 1024, 1024, (1, 1, 1), (2, 2, 2), (0, 0, 0)),      # layer4.0.downsample
    (3, 2, 2, 2, 256, 1024, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # layer4.x.conv3 (2x2x2 map: < one tile)
]


def _ext(i, k, s, p):
    return (i + 2 * p - k) // s + 1


@pytest.mark.parametrize('case', CASES)
def test_conv_fwd_dgrad_wgrad_tc(case):
    L = _lib()
    NB, T, H, W, Ci, Co, k, s, p = case
    taps = k[0] * k[1] * k[2]
    To, Ho, Wo = _ext(T, k[0], s[0], p[0]), _ext(H, k[1], s[1], p[1]), _ext(W, k[2], s[2], p[2])
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(NB, Ci, T, H, W, device='cuda', generator=g)
    w = torch.randn(Co, Ci, *k, device='cuda', generator=g) / math.sqrt(Ci * taps)
    wfh = torch.empty(Co, taps, Ci, dtype=torch.bfloat16, device='cuda')
    wfl = torch.empty_like(wfh)
    wdh = torch.empty(Ci, taps, Co, dtype=torch.bfloat16, device='cuda')
    wdl = torch.empty_like(wdh)
    L.pack_conv_weight_bf16(w.data_ptr(), wfh.data_ptr(), wfl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), Co, Ci, taps, _st())
    xh, xl = split(to_rows(x))
    geom = ConvGeom(NB, T, H, W, Ci, To, Ho, Wo, Co, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2])
    y = torch.full((NB * To * Ho * Wo, Co), float('nan'), device='cuda')
    ws = torch.empty(2 * Co, dtype=torch.float64, device='cuda')
    L.conv3d_fwd_tc(geom, xh.data_ptr(), xl.data_ptr(), wfh.data_ptr(), wfl.data_ptr(), y.data_ptr(), ws.data_ptr(), _st())
    torch.cuda.synchronize()
    # fused BatchNorm statistics of the conv output
    mean, rstd = torch.empty(Co, device='cuda'), torch.empty(Co, device='cuda')
    L.bn_finalize(ws.data_ptr(), y.shape[0], Co, 1e-5, mean.data_ptr(), rstd.data_ptr(), _st())
    assert rel(mean, y.double().mean(0)) < 1e-5 or float((mean - y.mean(0)).abs().max()) < 1e-6
    assert rel(rstd, 1 / torch.sqrt(y.double().var(0, unbiased=False) + 1e-5)) < 1e-5
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yref = F.conv3d(xr, wr, None, s, p)
    assert tuple(yref.shape[2:]) == (To, Ho, Wo)
    assert not torch.isnan(y).any()
    assert rel(from_rows(y, NB, To, Ho, Wo), yref) < 5e-5
    dy = torch.randn(yref.shape, device='cuda', generator=g)
    yref.backward(dy)
    dh, dl = split(to_rows(dy))
    # dgrad (accumulating into a non-zero base: the 1x1 strided sites leave the other parity classes untouched)
    base = torch.randn(NB * T * H * W, Ci, device='cuda', generator=g)
    dx = base.clone()
    L.conv3d_dgrad_tc(geom, dh.data_ptr(), dl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), dx.data_ptr(), 1, _st())
    torch.cuda.synchronize()
    assert rel(dx - base, to_rows(xr.grad)) < 5e-5
    if taps > 1 or s == (1, 1, 1):
        dx0 = torch.full_like(base, float('nan'))
        L.conv3d_dgrad_tc(geom, dh.data_ptr(), dl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), dx0.data_ptr(), 0, _st())
        assert not torch.isnan(dx0).any()
        assert rel(dx0, to_rows(xr.grad)) < 5e-5
    if s == (1, 1, 1):
        # dgrad with the BatchNorm-backward sums of the consumer BN fused into the epilogue (engine: bn1 of a block
        # from conv2's dgrad, the previous block's bn2 from conv1's accumulating dgrad)
        rows_in = NB * T * H * W
        ybn = torch.randn(rows_in, Ci, device='cuda', generator=g) * 2 + 0.3
        out_act = torch.randn(rows_in, Ci, device='cuda', generator=g)           # the BN's ReLU output: sign = mask
        out_hi = out_act.to(torch.bfloat16)
        bmean = torch.randn(Ci, device='cuda', generator=g) * 0.1
        brstd = torch.rand(Ci, device='cuda', generator=g) + 0.5
        for acc, use_mask in ((0, True), (1, True), (1, False)):
            dxf = base.clone() if acc else torch.full_like(base, float('nan'))
            wsf = torch.full((2 * Ci,), float('nan'), dtype=torch.float64, device='cuda')
            L.conv3d_dgrad_bnred_tc(geom, dh.data_ptr(), dl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), dxf.data_ptr(), acc,
                                    out_hi.data_ptr() if use_mask else None, ybn.data_ptr(), bmean.data_ptr(),
                                    brstd.data_ptr(), wsf.data_ptr(), _st())
            torch.cuda.synchronize()
            assert torch.equal(dxf, dx if acc else dx0)                            # same dx as the plain dgrad
            gm = dxf.double() * ((out_hi.float() > 0).double() if use_mask else 1.0)
            xhat = (ybn.double() - bmean.double()) * brstd.double()
            ref_ws = torch.cat([gm.sum(0), (gm * xhat).sum(0)])
            scale = float(torch.cat([gm.abs().sum(0), (gm * xhat).abs().sum(0)]).max())
            assert float((wsf - ref_ws).abs().max()) < 2e-6 * scale, (acc, use_mask)
    # wgrad
    dwp = torch.empty(Co, taps, Ci, device='cuda')
    dw = torch.full_like(w, float('nan'))
    L.conv3d_wgrad_tc(geom, xh.data_ptr(), xl.data_ptr(), dh.data_ptr(), dl.data_ptr(), dwp.data_ptr(), dw.data_ptr(), _st())
    torch.cuda.synchronize()
    assert not torch.isnan(dw).any()
    assert rel(dw, wr.grad) < 5e-5
