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
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def split(x):
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib().split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _st())
    return hi, lo


def to_rows(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().view(-1, x.shape[1])


def from_rows(r, NB, T, H, W):
    return r.view(NB, T, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


def test_split_bf16():
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(1 << 16, device='cuda', generator=g) * 3
    hi, lo = split(x)
    rec = hi.float() + lo.float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-20)).max()) < 2.0 ** -15
    assert torch.equal(hi, x.to(torch.bfloat16))


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (256, 256, 256), (1024, 768, 256), (200, 100, 128),
                                   (1617, 1617, 256), (6144, 512, 256)])
def test_gemm_nt_bf16x3(M, N, K):
    L = _lib()
    g = torch.Generator(device='cuda').manual_seed(2)
    A = torch.randn(M, K, device='cuda', generator=g)
    B = torch.randn(N, K, device='cuda', generator=g)
    ah, al = split(A)
    bh, bl = split(B)
    C = torch.full((M, N), float('nan'), device='cuda')
    L.gemm_nt_bf16x3_tc(M, N, K, ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), C.data_ptr(), _st())
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    assert not torch.isnan(C).any()
    assert rel(C, ref) < 5e-5


CASES = [
    # NB, T, H, W, Ci, Co, k, p
    (2, 5, 32, 32, 64, 64, (1, 3, 3), (0, 1, 1)),
    (3, 5, 16, 16, 128, 128, (1, 3, 3), (0, 1, 1)),
    (4, 3, 8, 8, 256, 256, (3, 3, 3), (1, 1, 1)),
    (8, 2, 4, 4, 256, 256, (3, 3, 3), (1, 1, 1)),
    (5, 2, 7, 7, 256, 256, (3, 3, 3), (1, 1, 1)),
    (3, 3, 14, 14, 256, 256, (3, 3, 3), (1, 1, 1)),
    (2, 5, 28, 28, 128, 128, (1, 3, 3), (0, 1, 1)),
    (3, 1, 8, 8, 64, 128, (1, 1, 1), (0, 0, 0)),
]


@pytest.mark.parametrize('case', CASES)
def test_conv_s1_fwd_and_dgrad(case):
    L = _lib()
    NB, T, H, W, Ci, Co, k, p = case
    taps = k[0] * k[1] * k[2]
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(NB, Ci, T, H, W, device='cuda', generator=g)
    w = torch.randn(Co, Ci, *k, device='cuda', generator=g) / math.sqrt(Ci * taps)
    wfh = torch.empty(Co, taps, Ci, dtype=torch.bfloat16, device='cuda')
    wfl, wdh, wdl = torch.empty_like(wfh), torch.empty(Ci, taps, Co, dtype=torch.bfloat16, device='cuda'), None
    wdl = torch.empty_like(wdh)
    L.pack_conv_weight_bf16(w.data_ptr(), wfh.data_ptr(), wfl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), Co, Ci, taps, _st())
    xh, xl = split(to_rows(x))
    geom = ConvGeom(NB, T, H, W, Ci, T, H, W, Co, k[0], k[1], k[2], 1, 1, 1, p[0], p[1], p[2])
    y = torch.full((NB * T * H * W, Co), float('nan'), device='cuda')
    L.conv3d_s1_tc(geom, xh.data_ptr(), xl.data_ptr(), wfh.data_ptr(), wfl.data_ptr(), y.data_ptr(), 0, _st())
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    yref = F.conv3d(xr, w, None, 1, p)
    assert not torch.isnan(y).any()
    assert rel(from_rows(y, NB, T, H, W), yref) < 5e-5
    # dgrad: same kernel on dy planes with the flipped/transposed weights
    dy = torch.randn(yref.shape, device='cuda', generator=g)
    yref.backward(dy)
    dh, dl = split(to_rows(dy))
    geom_d = ConvGeom(NB, T, H, W, Co, T, H, W, Ci, k[0], k[1], k[2], 1, 1, 1, p[0], p[1], p[2])
    base = torch.randn(NB * T * H * W, Ci, device='cuda', generator=g)
    dx = base.clone()
    L.conv3d_s1_tc(geom_d, dh.data_ptr(), dl.data_ptr(), wdh.data_ptr(), wdl.data_ptr(), dx.data_ptr(), 1, _st())
    torch.cuda.synchronize()
    assert rel(dx - base, to_rows(xr.grad)) < 5e-5
