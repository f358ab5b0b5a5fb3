"""Host-side orchestration of the DPC-RNN hot path over the C ABI (include/dpc_b200.h).

Two stages, each a forward/backward pair driven from a torch.autograd.Function
(dpc_b200/resnet_2d3d.py, dpc_b200/model_3d.py):

  backbone : ResNet2d3d_full.forward            (/root/reference/backbone/resnet_2d3d.py:259-270)
  head     : avg-pool/ReLU split, ConvGRU, predictor loop, score matmul
             (/root/reference/dpc/model_3d.py:53-83, backbone/convrnn.py:24-34,62-88)

Data layout in HBM: activations are channels-last rows [NB*T*H*W, C] fp32 (+ split-bf16 hi/lo planes
for tensor-core operands); the caller's NCDHW video block is repacked once per step into 2x2
space-to-depth split-bf16 planes that the stem forward and wgrad read through TMA.  PyTorch provides
memory (caching allocator) and the current stream only; every FLOP below is issued through
libdpc_b200.so.
"""
import math
import os

import torch

from ._lib import lib, ptr, ConvGeom
from .arch import backbone_spec, FEATURE_SIZE

BN_EPS = 1e-5

# optional per-kernel-family device timer (bench.py's roofline leg); None on the normal path
_TIMER = None


def set_timer(timer):
    """timer: None or an object with .start(tag) -> token and .stop(token) using CUDA events"""
    global _TIMER
    _TIMER = timer


class EventTimer:
    """CUDA-event timer on the current stream; .totals() -> {tag: (calls, ms)} after a synchronize"""

    def __init__(self, keep_overlap=False):
        self.records = []
        self.streams = []                 # cuda stream handle of every record (timeline diagnostics)
        self.keep_overlap = keep_overlap  # True: the backward keeps its chain / side streams while being timed

    def start(self, tag, detail=None):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        return (tag, a, b, detail)

    def stop(self, tok):
        tok[2].record()
        self.records.append(tok)
        self.streams.append(torch.cuda.current_stream().cuda_stream)

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for tag, a, b, _ in self.records:
            c, t = out.get(tag, (0, 0.0))
            out[tag] = (c + 1, t + a.elapsed_time(b))
        return out

    def calls(self, tag):
        """[(detail, ms)] of every call recorded under `tag`; detail = (Ci, Co, taps, rows_out, stride-1?) for conv sites"""
        torch.cuda.synchronize()
        return [(d, a.elapsed_time(b)) for t, a, b, d in self.records if t == tag]


def _timed(tag):
    def deco(fn):
        def wrapped(*a, **k):
            if _TIMER is None:
                return fn(*a, **k)
            site = a[0] if a and hasattr(a[0], 'geom') else None
            detail = None
            if site is not None:
                g = site.geom
                detail = (site.Ci, site.Co, site.taps, site.rows_out, g.sT * g.sH * g.sW == 1)
            tok = _TIMER.start(tag, detail)
            try:
                return fn(*a, **k)
            finally:
                _TIMER.stop(tok)
        return wrapped
    return deco


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _empty(shape, ref, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=ref.device)


def _out_extent(i, k, s, p):
    return (i + 2 * p - k) // s + 1


class ConvSite:
    """geometry of one Conv3d call site"""
    __slots__ = ('geom', 'taps', 'Ci', 'Co', 'rows_in', 'rows_out', 'dims_out')

    def __init__(self, NB, dims_in, Ci, Co, k, s, p):
        Ti, Hi, Wi = dims_in
        To, Ho, Wo = (_out_extent(Ti, k[0], s[0], p[0]), _out_extent(Hi, k[1], s[1], p[1]),
                      _out_extent(Wi, k[2], s[2], p[2]))
        self.geom = ConvGeom(NB, Ti, Hi, Wi, Ci, To, Ho, Wo, Co, k[0], k[1], k[2], s[0], s[1], s[2],
                             p[0], p[1], p[2])
        self.taps = k[0] * k[1] * k[2]
        self.Ci, self.Co = Ci, Co
        self.rows_in = NB * Ti * Hi * Wi
        self.rows_out = NB * To * Ho * Wo
        self.dims_out = (To, Ho, Wo)


# BatchNorm-backward reductions in the epilogue of the stride-1 dgrads that produce their input (dgrad_bnred): False,
# 'halo' (only the 64 -> 64 1x3x3 sites of layer1, whose halo-patch kernel prefetches the mask / y rows before it waits
# for the accumulators) or True (every stride-1 site; measured slower on the tap-per-box kernels at B = 128: the per-row
# gathers stall their 4-warp epilogue: layer2 +0.38 ms vs a 0.26 ms reduce pass, layer3 +0.31 vs 0.10).
FUSE_BN_REDUCE = False
# conv1's wgrad inside the recomputing stem backward kernel (stem_pool.cu); False = gradient planes to HBM + the separate
# space-to-depth wgrad kernel (also the automatic fallback when the fused schedule does not fit shared memory)
STEM_FUSE_WGRAD = True


# Plane dtype of conv operands.  tcgen05 kind::f16 needs ONE input format per instruction and wgrad multiplies activations by
# gradients (which need bf16's exponent range), so activations, weights and gradients are all bf16 pairs; only the score
# matmul -- both operands forward values of O(1) magnitude -- runs on fp16 pairs (f16=True below).
ACT = GRD = torch.bfloat16


@_timed('split_bf16')
def _split(x, st, f16=False):
    """fp32 rows -> (hi, lo) 16-bit planes with hi + lo ~= x: bf16 pairs (~2^-17, fp32 range), or fp16 pairs (~2^-22) for
    forward values that only ever meet other forward values (f16=True)"""
    dt = torch.float16 if f16 else torch.bfloat16
    hi = torch.empty(x.shape, dtype=dt, device=x.device)
    lo = torch.empty(x.shape, dtype=dt, device=x.device)
    (lib().split_f16 if f16 else lib().split_bf16)(ptr(x), ptr(hi), ptr(lo), x.numel(), st)
    return hi, lo


class TcConvSite(ConvSite):
    """Conv3d call site on the tcgen05 kernels; activations/gradients are passed as (hi, lo) planes"""
    __slots__ = ('wfh', 'wfl', 'wdh', 'wdl')

    def pack(self, w, st):
        self.wfh = torch.empty((self.Co, self.taps, self.Ci), dtype=ACT, device=w.device)
        self.wfl = torch.empty((self.Co, self.taps, self.Ci), dtype=ACT, device=w.device)
        self.wdh = torch.empty((self.Ci, self.taps, self.Co), dtype=GRD, device=w.device)
        self.wdl = torch.empty((self.Ci, self.taps, self.Co), dtype=GRD, device=w.device)
        lib().pack_conv_weight_bf16(ptr(w), ptr(self.wfh), ptr(self.wfl), ptr(self.wdh), ptr(self.wdl),
                                    self.Co, self.Ci, self.taps, st)

    @_timed('conv_fwd')
    def fwd(self, xp, st):
        y = torch.empty((self.rows_out, self.Co), dtype=torch.float32, device=xp[0].device)
        lib().conv3d_fwd_tc(self.geom, ptr(xp[0]), ptr(xp[1]), ptr(self.wfh), ptr(self.wfl), ptr(y), None, st)
        return y

    @_timed('conv_fwd')
    def fwd_bn(self, xp, st):
        """conv + fused BatchNorm batch statistics -> (y, mean, rstd)"""
        dev = xp[0].device
        y = torch.empty((self.rows_out, self.Co), dtype=torch.float32, device=dev)
        ws = torch.empty(2 * self.Co, dtype=torch.float64, device=dev)
        mean = torch.empty(self.Co, dtype=torch.float32, device=dev)
        rstd = torch.empty(self.Co, dtype=torch.float32, device=dev)
        L = lib()
        L.conv3d_fwd_tc(self.geom, ptr(xp[0]), ptr(xp[1]), ptr(self.wfh), ptr(self.wfl), ptr(y), ptr(ws), st)
        L.bn_finalize(ptr(ws), self.rows_out, self.Co, BN_EPS, ptr(mean), ptr(rstd), st)
        return y, mean, rstd

    @_timed('conv_dgrad')
    def dgrad(self, dyp, st, dx=None):
        acc = 1
        if dx is None:
            dx = torch.empty((self.rows_in, self.Ci), dtype=torch.float32, device=dyp[0].device)
            acc = 0
        lib().conv3d_dgrad_tc(self.geom, ptr(dyp[0]), ptr(dyp[1]), ptr(self.wdh), ptr(self.wdl), ptr(dx), acc, st)
        return dx

    @_timed('conv_dgrad')
    def dgrad_bnred(self, dyp, st, mask_hi, y, mean, rstd, dx=None):
        """stride-1 dgrad + the BatchNorm-backward reduction of the BN that consumes dx (mask_hi: hi plane of that
        BN's ReLU output or None, y / mean / rstd: its input and statistics) -> (dx, ws [2*Ci] float64)"""
        acc = 1
        if dx is None:
            dx = torch.empty((self.rows_in, self.Ci), dtype=torch.float32, device=dyp[0].device)
            acc = 0
        ws = torch.empty(2 * self.Ci, dtype=torch.float64, device=dx.device)
        lib().conv3d_dgrad_bnred_tc(self.geom, ptr(dyp[0]), ptr(dyp[1]), ptr(self.wdh), ptr(self.wdl), ptr(dx), acc,
                                    ptr(mask_hi), ptr(y), ptr(mean), ptr(rstd), ptr(ws), st)
        return dx, ws

    @property
    def stride1(self):
        g = self.geom
        return g.sT == 1 and g.sH == 1 and g.sW == 1

    @property
    def fuse_bnred(self):
        """whether this site's dgrad should reduce the consumer BN's backward sums in its epilogue (FUSE_BN_REDUCE)"""
        if not FUSE_BN_REDUCE or not self.stride1:
            return False
        g = self.geom
        halo = (g.kT, g.kH, g.kW) == (1, 3, 3) and self.Ci == 64 and self.Co == 64
        return halo or FUSE_BN_REDUCE is True

    @_timed('conv_wgrad')
    def wgrad(self, xp, dyp, st):
        g = self.geom
        dev = xp[0].device
        dwp = torch.empty((self.Co, self.taps, self.Ci), dtype=torch.float32, device=dev)
        dw = torch.empty((self.Co, self.Ci, g.kT, g.kH, g.kW), dtype=torch.float32, device=dev)
        lib().conv3d_wgrad_tc(g, ptr(xp[0]), ptr(xp[1]), ptr(dyp[0]), ptr(dyp[1]), ptr(dwp), ptr(dw), st)
        return dw


@_timed('bn_stats')
def _bn_stats(y, rows, C, st):
    ws = torch.empty(2 * C, dtype=torch.float64, device=y.device)
    mean = _empty((C,), y)
    rstd = _empty((C,), y)
    lib().bn_stats(ptr(y), rows, C, ptr(ws), ptr(mean), ptr(rstd), BN_EPS, st)
    return mean, rstd


@_timed('bn_apply')
def _bn_apply(y, mean, rstd, gamma, beta, relu, rows, C, st, res=None, res_planes=None, rbn=None,
              want_rows=True, want_planes=False):
    """[relu](bn(y) + residual) -> (fp32 rows or None, (hi, lo) planes or None)"""
    out = torch.empty_like(y) if want_rows else None
    planes = None
    if want_planes:
        planes = (torch.empty(y.shape, dtype=ACT, device=y.device), torch.empty(y.shape, dtype=ACT, device=y.device))
    r = rbn if rbn is not None else (None, None, None, None)
    rp = res_planes if (res is None and res_planes is not None) else (None, None)
    lib().bn_apply_fwd(ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(res), ptr(rp[0]), ptr(rp[1]),
                       ptr(r[0]), ptr(r[1]), ptr(r[2]), ptr(r[3]), 1 if relu else 0, ptr(out),
                       ptr(planes[0]) if planes else None, ptr(planes[1]) if planes else None, rows, C, st)
    return out, planes


@_timed('bn_bwd')
def _bn_bwd(dout, out, relu, y, mean, rstd, gamma, rows, C, st, want_g=False, out_hi=None, want_rows=True,
            want_planes=False, ws=None, frozen=False):
    """-> (dy rows or None, dy planes or None, dgamma, dbeta, g or None).  ws: the reduction [sum g | sum g*xhat]
    when the dgrad that produced `dout` already computed it in its epilogue (TcConvSite.dgrad_bnred).
    frozen: the BN normalised with FIXED statistics (eval mode of track_running_stats=True): dgamma / dbeta as usual,
    but dy = gamma * rstd * g without the batch-statistics terms (a rarely used path: reduce, then an apply pass over
    zeroed sums)"""
    if frozen:
        res = _bn_bwd(dout, out, relu, y, mean, rstd, gamma, rows, C, st, want_g=want_g, out_hi=out_hi, want_rows=want_rows,
                      want_planes=want_planes)
        zero = torch.zeros(2 * C, dtype=torch.float64, device=y.device)
        dy, planes, _, _, g = _bn_bwd(dout, out, relu, y, mean, rstd, gamma, rows, C, st, want_g=want_g, out_hi=out_hi,
                                      want_rows=want_rows, want_planes=want_planes, ws=zero)
        return dy, planes, res[2], res[3], g
    fused = ws is not None
    if not fused:
        ws = torch.empty(2 * C, dtype=torch.float64, device=y.device)
    dgamma = _empty((C,), y)
    dbeta = _empty((C,), y)
    dy = torch.empty_like(y) if want_rows else None
    planes = None
    if want_planes:
        planes = (torch.empty(y.shape, dtype=GRD, device=y.device), torch.empty(y.shape, dtype=GRD, device=y.device))
    g = torch.empty_like(y) if want_g else None
    fn = lib().bn_bwd_apply if fused else lib().bn_bwd
    fn(ptr(dout), ptr(out), ptr(out_hi) if out is None else None, 1 if relu else 0, ptr(y), ptr(mean),
       ptr(rstd), ptr(gamma), ptr(ws), ptr(dgamma), ptr(dbeta), ptr(dy),
       ptr(planes[0]) if planes else None, ptr(planes[1]) if planes else None, ptr(g), rows, C, st)
    return dy, planes, dgamma, dbeta, g


# ==================================================================================================
# backbone
# ==================================================================================================
def backbone_param_names(network):
    names = ['conv1.weight', 'bn1.weight', 'bn1.bias']
    for b in backbone_spec(network):
        p = b['name']
        names += [p + '.conv1.weight', p + '.bn1.weight', p + '.bn1.bias',
                  p + '.conv2.weight', p + '.bn2.weight', p + '.bn2.bias']
        if b['block'] == 'bottleneck':
            names += [p + '.conv3.weight', p + '.bn3.weight', p + '.bn3.bias']
        if b['downsample']:
            names += [p + '.downsample.0.weight', p + '.downsample.1.weight', p + '.downsample.1.bias']
    return names


BN_MOMENTUM = 0.1


def _conv_bn(site, op, name, bn_state, training, st):
    """conv + the BatchNorm statistics to normalise with: batch statistics (and, with running buffers in train
    mode, their momentum update) or the running buffers (eval mode of track_running_stats=True)."""
    L = lib()
    if bn_state is not None and not training:
        y = site.fwd(op, st)
        rm, rv = bn_state[name]
        rstd = torch.empty_like(rm)
        L.bn_rstd_from_var(ptr(rv), BN_EPS, ptr(rstd), rm.numel(), st)
        return y, rm, rstd
    y, m, r = site.fwd_bn(op, st)
    if bn_state is not None:
        rm, rv = bn_state[name]
        L.bn_running_update(ptr(m), ptr(r), site.rows_out, BN_EPS, BN_MOMENTUM, ptr(rm), ptr(rv), rm.numel(), st)
    return y, m, r


def _stem_forward_unpooled(L, st, x, P, bn_state, training, running, network):
    """conv1 -> y0 (stored) -> bn1 + relu + maxpool for frames the pooled-stem schedule does not cover (odd H / W):
    stem_tc.cu builds the im2col tile in shared memory.  Returns (a0 rows, conv operand planes, ctx)."""
    NB, Cin, T, H, W = x.shape
    Ho, Wo = _out_extent(H, 7, 2, 3), _out_extent(W, 7, 2, 3)
    rows0 = NB * T * Ho * Wo
    y0 = _empty((rows0, 64), x)
    ws0 = None if running else torch.empty(128, dtype=torch.float64, device=x.device)
    _timed('stem_fwd')(L.stem_conv_fwd_tc)(ptr(x), ptr(P['conv1.weight']), ptr(y0), ptr(ws0), NB, T, H, W, st)
    if running:
        m0 = bn_state['bn1'][0]
        r0 = torch.empty_like(m0)
        L.bn_rstd_from_var(ptr(bn_state['bn1'][1]), BN_EPS, ptr(r0), 64, st)
    else:
        m0, r0 = _empty((64,), x), _empty((64,), x)
        L.bn_finalize(ptr(ws0), rows0, 64, BN_EPS, ptr(m0), ptr(r0), st)
        if bn_state is not None:
            L.bn_running_update(ptr(m0), ptr(r0), rows0, BN_EPS, BN_MOMENTUM, ptr(bn_state['bn1'][0]), ptr(bn_state['bn1'][1]), 64, st)
    Hp, Wp = _out_extent(Ho, 3, 2, 1), _out_extent(Wo, 3, 2, 1)
    a0 = _empty((NB * T * Hp * Wp, 64), x)
    _timed('stem_pool_fwd')(L.bn_relu_maxpool_fwd)(ptr(y0), ptr(m0), ptr(r0), ptr(P['bn1.weight']), ptr(P['bn1.bias']), ptr(a0),
                                                   NB * T, Ho, Wo, 64, st)
    ctx = dict(network=network, x=x, y0=y0, m0=m0, r0=r0, a0=a0, stem_dims=(NB, T, H, W, Ho, Wo), blocks=[])
    return a0, _split(a0, st), ctx


def backbone_forward(network, x, P, need_ctx=True, bn_state=None, training=True):
    """x [NB,3,T,H,W] fp32 contiguous CUDA; P: name -> tensor.  Returns (feature rows [NB*To*Ho*Wo, 256],
    (To,Ho,Wo), ctx).  bn_state: None (track_running_stats=False: batch statistics always) or
    {bn name: (running_mean, running_var)} (track_running_stats=True, eval/model_3d_lc.py:26)."""
    L = lib()
    st = _stream()
    running = bn_state is not None and not training
    NB, Cin, T, H, W = x.shape
    if Cin != 3:
        raise ValueError('backbone expects 3 input channels, got %d' % Cin)
    Ho, Wo = _out_extent(H, 7, 2, 3), _out_extent(W, 7, 2, 3)
    rows0 = NB * T * Ho * Wo
    x2 = None
    Hp, Wp = _out_extent(Ho, 3, 2, 1), _out_extent(Wo, 3, 2, 1)
    rows_p = NB * T * Hp * Wp
    bf = dict(dtype=ACT, device=x.device)
    pooled = L.stem_pool_supported(H, W) >= 1
    if pooled:
        # conv1 + bn1 + relu + maxpool with the conv1 output never stored (stem_pool.cu): the kernel keeps, per pooled
        # position, the conv value the pool selects (+ its window index) and bn1's batch sums over all conv positions
        x2 = (torch.empty((NB, T, H // 2, W // 2, 16), **bf), torch.empty((NB, T, H // 2, W // 2, 16), **bf))
        wp = torch.empty(32768, **bf)
        ypool = _empty((rows_p, 64), x)
        idx = torch.empty((rows_p, 64), dtype=torch.uint8, device=x.device)
        ws0 = None if running else torch.empty(128, dtype=torch.float64, device=x.device)
        _timed('stem_fwd')(L.stem_s2d_pack)(ptr(x), ptr(x2[0]), ptr(x2[1]), NB, T, H, W, st)
        L.stem_s2d_wpack(ptr(P['conv1.weight']), ptr(wp), st)
        _timed('stem_fwd')(L.stem_pool_fwd)(ptr(x2[0]), ptr(x2[1]), ptr(wp), ptr(P['bn1.weight']), ptr(ypool), ptr(idx), ptr(ws0),
                                            NB, T, H, W, st)
        if running:
            m0 = bn_state['bn1'][0]
            r0 = torch.empty_like(m0)
            L.bn_rstd_from_var(ptr(bn_state['bn1'][1]), BN_EPS, ptr(r0), 64, st)
        else:
            m0, r0 = _empty((64,), x), _empty((64,), x)
            L.bn_finalize(ptr(ws0), rows0, 64, BN_EPS, ptr(m0), ptr(r0), st)
            if bn_state is not None:
                L.bn_running_update(ptr(m0), ptr(r0), rows0, BN_EPS, BN_MOMENTUM, ptr(bn_state['bn1'][0]), ptr(bn_state['bn1'][1]), 64, st)
        a0p = (torch.empty((rows_p, 64), **bf), torch.empty((rows_p, 64), **bf))
        _timed('stem_pool_fwd')(L.stem_pool_finalize)(ptr(ypool), ptr(idx), ptr(m0), ptr(r0), ptr(P['bn1.weight']), ptr(P['bn1.bias']),
                                                      ptr(a0p[0]), ptr(a0p[1]), None, rows_p, st)
        ctx = dict(network=network, x=x, x2=x2, wp=wp, ypool=ypool, idx=idx, m0=m0, r0=r0, stem_dims=(NB, T, H, W, Ho, Wo), blocks=[],
                   frozen=running)
        cur, cur_op = None, a0p
    else:
        cur, cur_op, ctx = _stem_forward_unpooled(L, st, x, P, bn_state, training, running, network)
        ctx['frozen'] = running
    dims, C = (T, Hp, Wp), 64
    tc = True                      # conv operands are always split-bf16 planes (tcgen05 kernels)
    Site = TcConvSite
    spec = backbone_spec(network)
    for bi, b in enumerate(spec):
        p = b['name']
        last = bi + 1 == len(spec)
        if b['is3d']:
            k, pad = (3, 3, 3), (1, 1, 1)
            s1 = (b['stride'],) * 3
        else:
            k, pad = (1, 3, 3), (0, 1, 1)
            s1 = (1, b['stride'], b['stride'])
        want_rows = (not tc) or last                 # the last block's output feeds the head as fp32 rows
        want_planes = tc and not last
        if b['block'] == 'bottleneck':
            # Bottleneck2d / 3d (resnet_2d3d.py:119-202): 1x1x1 -> k (carries the stride) -> 1x1x1 (x4 channels)
            one, nopad = (1, 1, 1), (0, 0, 0)
            c1 = Site(NB, dims, b['inplanes'], b['planes'], one, one, nopad)
            c1.pack(P[p + '.conv1.weight'], st)
            y1, m1, r1 = _conv_bn(c1, cur_op, p + '.bn1', bn_state, training, st)
            a1, a1_pl = _bn_apply(y1, m1, r1, P[p + '.bn1.weight'], P[p + '.bn1.bias'], True, c1.rows_out, c1.Co, st,
                                  want_rows=not tc, want_planes=tc)
            a1_op = a1_pl if tc else a1
            c2 = Site(NB, c1.dims_out, b['planes'], b['planes'], k, s1, pad)
            c2.pack(P[p + '.conv2.weight'], st)
            y2, m2, r2 = _conv_bn(c2, a1_op, p + '.bn2', bn_state, training, st)
            a2, a2_pl = _bn_apply(y2, m2, r2, P[p + '.bn2.weight'], P[p + '.bn2.bias'], True, c2.rows_out, c2.Co, st,
                                  want_rows=not tc, want_planes=tc)
            a2_op = a2_pl if tc else a2
            c3 = Site(NB, c2.dims_out, b['planes'], b['outplanes'], one, one, nopad)
            c3.pack(P[p + '.conv3.weight'], st)
            y3, m3, r3 = _conv_bn(c3, a2_op, p + '.bn3', bn_state, training, st)
            rec = dict(spec=b, c1=c1, c2=c2, c3=c3, xin_op=cur_op, y1=y1, m1=m1, r1=r1, a1=a1, a1_op=a1_op,
                       y2=y2, m2=m2, r2=r2, a2=a2, a2_op=a2_op, y3=y3, m3=m3, r3=r3, tc=tc)
            if b['downsample']:
                cd = Site(NB, dims, b['inplanes'], b['outplanes'], one, s1, nopad)
                cd.pack(P[p + '.downsample.0.weight'], st)
                yd, md, rd = _conv_bn(cd, cur_op, p + '.downsample.1', bn_state, training, st)
                out, out_pl = _bn_apply(y3, m3, r3, P[p + '.bn3.weight'], P[p + '.bn3.bias'], b['final_relu'],
                                        c3.rows_out, c3.Co, st, res=yd,
                                        rbn=(md, rd, P[p + '.downsample.1.weight'], P[p + '.downsample.1.bias']),
                                        want_rows=want_rows, want_planes=want_planes)
                rec.update(cd=cd, yd=yd, md=md, rd=rd)
            else:
                out, out_pl = _bn_apply(y3, m3, r3, P[p + '.bn3.weight'], P[p + '.bn3.bias'], b['final_relu'],
                                        c3.rows_out, c3.Co, st, res=cur, res_planes=cur_op if tc else None,
                                        want_rows=want_rows, want_planes=want_planes)
            rec['out'], rec['out_hi'] = out, (out_pl[0] if out_pl else None)
            if need_ctx:
                ctx['blocks'].append(rec)
            cur, dims, C = out, c3.dims_out, b['outplanes']
            cur_op = out_pl if tc else out
            continue
        c1 = Site(NB, dims, b['inplanes'], b['planes'], k, s1, pad)
        c1.pack(P[p + '.conv1.weight'], st)
        y1, m1, r1 = _conv_bn(c1, cur_op, p + '.bn1', bn_state, training, st)
        # tensor-core path: the normalise pass writes the next conv's operand planes directly
        a1, a1_pl = _bn_apply(y1, m1, r1, P[p + '.bn1.weight'], P[p + '.bn1.bias'], True, c1.rows_out, c1.Co, st,
                              want_rows=not tc, want_planes=tc)
        a1_op = a1_pl if tc else a1
        c2 = Site(NB, c1.dims_out, b['planes'], b['planes'], k, (1, 1, 1), pad)
        c2.pack(P[p + '.conv2.weight'], st)
        y2, m2, r2 = _conv_bn(c2, a1_op, p + '.bn2', bn_state, training, st)
        rec = dict(spec=b, c1=c1, c2=c2, xin_op=cur_op, y1=y1, m1=m1, r1=r1, a1=a1, a1_op=a1_op,
                   y2=y2, m2=m2, r2=r2, tc=tc)
        if b['downsample']:
            cd = Site(NB, dims, b['inplanes'], b['planes'], (1, 1, 1), s1, (0, 0, 0))
            cd.pack(P[p + '.downsample.0.weight'], st)
            yd, md, rd = _conv_bn(cd, cur_op, p + '.downsample.1', bn_state, training, st)
            out, out_pl = _bn_apply(y2, m2, r2, P[p + '.bn2.weight'], P[p + '.bn2.bias'], b['final_relu'],
                                    c2.rows_out, c2.Co, st, res=yd,
                                    rbn=(md, rd, P[p + '.downsample.1.weight'], P[p + '.downsample.1.bias']),
                                    want_rows=want_rows, want_planes=want_planes)
            rec.update(cd=cd, yd=yd, md=md, rd=rd)
        else:
            out, out_pl = _bn_apply(y2, m2, r2, P[p + '.bn2.weight'], P[p + '.bn2.bias'], b['final_relu'],
                                    c2.rows_out, c2.Co, st, res=cur, res_planes=cur_op if tc else None,
                                    want_rows=want_rows, want_planes=want_planes)
        rec['out'], rec['out_hi'] = out, (out_pl[0] if out_pl else None)
        if need_ctx:
            ctx['blocks'].append(rec)
        cur, dims, C = out, c2.dims_out, b['planes']
        cur_op = out_pl if tc else out
    return cur, dims, (ctx if need_ctx else None)


# Weight gradients are leaves of the backward graph: nothing on the dgrad -> BN-backward chain waits for
# them.  They run on a side stream so the tensor-bound wgrad kernels overlap the HBM-bound BN-backward
# passes of the following layers (one side stream per device, joined before backward returns).
_SIDE_STREAMS = {}
OVERLAP_WGRAD = True


def _side_stream(device):
    s = _SIDE_STREAMS.get(device)
    if s is None:
        s = _SIDE_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


def _wgrad_async(site, x_op, dy_op, main, side):
    """site.wgrad(x_op, dy_op) on the side stream, ordered after everything issued so far on `main`"""
    if side is None:
        return site.wgrad(x_op, dy_op, main.cuda_stream)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        dw = site.wgrad(x_op, dy_op, side.cuda_stream)
    for t in (x_op if isinstance(x_op, tuple) else (x_op,)) + (dy_op if isinstance(dy_op, tuple) else (dy_op,)):
        t.record_stream(side)          # the caching allocator must not recycle these before the side stream is done
    return dw


# When the weight gradient of a site is launched.  'early' (DPC_WGRAD_DEFER=0): as soon as its dy planes exist, i.e.
# concurrently with the dgrad of the same site -- two tensor-bound kernels that only take turns on the tensor pipe.
# 'deferred' (default): after that dgrad has been enqueued, so the wgrad starts under the HBM-bound BatchNorm-backward passes
# of the NEXT site.  Measured with scripts/timeline_step.py (B = 128, same box): 70.98 vs 71.91 ms/step.  Either way the
# overlap buys little: kernels that share the GPU slow each other by 40-60 % (bn_bwd 10.6 -> 15.3 ms, dgrad 15.6 -> 23 ms under
# 40 ms of concurrent wgrad) -- the step runs under `sw_power_cap` throughout, so concurrency trades clock for occupancy.
WGRAD_DEFER = os.environ.get('DPC_WGRAD_DEFER', '1') != '0'


class _WgradQueue:
    def __init__(self, main, side):
        self.main, self.side, self.q = main, side, []

    def add(self, G, key, site, x_op, dy_op):
        if self.side is None or not WGRAD_DEFER:
            G[key] = _wgrad_async(site, x_op, dy_op, self.main, self.side)
        else:
            self.q.append((G, key, site, x_op, dy_op))

    def flush(self):
        """call after the dgrad that consumes the same dy has been enqueued on the chain stream"""
        for G, key, site, x_op, dy_op in self.q:
            G[key] = _wgrad_async(site, x_op, dy_op, self.main, self.side)
        self.q.clear()


_CHAIN_STREAMS = {}


def _chain_stream(device):
    """high-priority stream for the dgrad -> BN-backward critical chain (the wgrad side stream is low priority,
    so its CTAs only fill SMs the chain leaves idle, e.g. under the HBM-bound BN passes)"""
    s = _CHAIN_STREAMS.get(device)
    if s is None:
        s = _CHAIN_STREAMS[device] = torch.cuda.Stream(device=device, priority=-1)
    return s


def backbone_backward(ctx, dout, P):
    """dout: rows [NB*To*Ho*Wo, 256].  Returns dict name -> grad (parameter layouts)."""
    if not (OVERLAP_WGRAD and (_TIMER is None or _TIMER.keep_overlap)):
        return _backbone_backward(ctx, dout, P, None)
    caller = torch.cuda.current_stream()
    chain = _chain_stream(dout.device)
    side = _side_stream(dout.device)
    # Stream invariant: every kernel below runs on `chain` (dgrad / BatchNorm backward) or `side` (weight gradients); both
    # are ordered after the caller's stream here and joined back into it before returning, so the caller may use G (and
    # free the saved activations) in plain stream order.  The gradients were allocated while `chain` / `side` were current:
    # record the caller's stream on them so the caching allocator does not recycle their memory early.
    chain.wait_stream(caller)
    with torch.cuda.stream(chain):
        G = _backbone_backward(ctx, dout, P, side)
    dout.record_stream(chain)
    caller.wait_stream(chain)
    for t in G.values():
        t.record_stream(caller)
    return G


def _stem_backward_unpooled(L, st, ctx, dout, P, G):
    """backward of _stem_forward_unpooled; fills G['bn1.*'] and returns dW of conv1"""
    NB, T, H, W, Ho, Wo = ctx['stem_dims']
    # max-pool bwd + ReLU bwd + bn1 bwd fused (two passes, no materialised pooling gradient)
    y0 = ctx['y0']
    dy0p = (torch.empty(y0.shape, dtype=GRD, device=y0.device), torch.empty(y0.shape, dtype=GRD, device=y0.device))
    ws = torch.empty(128, dtype=torch.float64, device=y0.device)
    G['bn1.weight'], G['bn1.bias'] = _empty((64,), y0), _empty((64,), y0)
    _timed('stem_tail_bwd')(L.stem_tail_bwd)(ptr(y0), ptr(ctx['m0']), ptr(ctx['r0']), ptr(P['bn1.weight']),
                                             ptr(P['bn1.bias']), ptr(ctx['a0']), ptr(dout), ptr(ws),
                                             ptr(G['bn1.weight']), ptr(G['bn1.bias']), None, ptr(dy0p[0]),
                                             ptr(dy0p[1]), NB * T, Ho, Wo, 64, 1, st)
    dw0 = torch.empty_like(P['conv1.weight'])
    _timed('stem_wgrad')(L.stem_conv_wgrad_tc)(ptr(ctx['x']), ptr(dy0p[0]), ptr(dy0p[1]), ptr(dw0), NB, T, H, W, st)
    return dw0


def _backbone_backward(ctx, dout, P, side):
    L = lib()
    st = _stream()
    main = torch.cuda.current_stream()
    wq = _WgradQueue(main, side)
    G = {}
    blocks = ctx['blocks']
    ws_out = None          # BN-backward sums of this block's bn2, if the dgrad that produced `dout` fused them
    for bi in reversed(range(len(blocks))):
        rec = blocks[bi]
        b = rec['spec']
        p = b['name']
        c1, c2 = rec['c1'], rec['c2']
        relu = b['final_relu']
        has_ds = b['downsample']
        tc = rec['tc']
        kw = dict(want_rows=not tc, want_planes=tc, frozen=ctx.get('frozen', False))          # conv-operand format of the dy tensors
        op = (lambda rows, planes: planes) if tc else (lambda rows, planes: rows)
        if b['block'] == 'bottleneck':
            c3 = rec['c3']
            dy3r, dy3p, G[p + '.bn3.weight'], G[p + '.bn3.bias'], g = _bn_bwd(
                dout, rec['out'], relu, rec['y3'], rec['m3'], rec['r3'], P[p + '.bn3.weight'],
                c3.rows_out, c3.Co, st, want_g=not has_ds, out_hi=rec['out_hi'], **kw)
            dy3 = op(dy3r, dy3p)
            if has_ds:
                cd = rec['cd']
                dydr, dydp, G[p + '.downsample.1.weight'], G[p + '.downsample.1.bias'], _ = _bn_bwd(
                    dout, rec['out'], relu, rec['yd'], rec['md'], rec['rd'], P[p + '.downsample.1.weight'],
                    cd.rows_out, cd.Co, st, out_hi=rec['out_hi'], **kw)
                dyd = op(dydr, dydp)
            del dout
            wq.add(G, p + '.conv3.weight', c3, rec['a2_op'], dy3)
            da2 = c3.dgrad(dy3, st)
            wq.flush()
            del dy3, dy3r, dy3p
            dy2r, dy2p, G[p + '.bn2.weight'], G[p + '.bn2.bias'], _ = _bn_bwd(
                da2, rec['a2'], True, rec['y2'], rec['m2'], rec['r2'], P[p + '.bn2.weight'],
                c2.rows_out, c2.Co, st, out_hi=(rec['a2_op'][0] if tc else None), **kw)
            dy2 = op(dy2r, dy2p)
            del da2
            wq.add(G, p + '.conv2.weight', c2, rec['a1_op'], dy2)
            da1 = c2.dgrad(dy2, st)
            wq.flush()
            del dy2, dy2r, dy2p
            dy1r, dy1p, G[p + '.bn1.weight'], G[p + '.bn1.bias'], _ = _bn_bwd(
                da1, rec['a1'], True, rec['y1'], rec['m1'], rec['r1'], P[p + '.bn1.weight'],
                c1.rows_out, c1.Co, st, out_hi=(rec['a1_op'][0] if tc else None), **kw)
            dy1 = op(dy1r, dy1p)
            del da1
            wq.add(G, p + '.conv1.weight', c1, rec['xin_op'], dy1)
            if has_ds:
                dx = c1.dgrad(dy1, st)
                cd.dgrad(dyd, st, dx=dx)
                wq.add(G, p + '.downsample.0.weight', cd, rec['xin_op'], dyd)
                wq.flush()
                del dyd, dydr, dydp
            else:
                dx = c1.dgrad(dy1, st, dx=g)          # dx = g + dgrad
                wq.flush()
            del dy1, dy1r, dy1p
            dout = dx
            rec.clear()
            continue
        dy2r, dy2p, G[p + '.bn2.weight'], G[p + '.bn2.bias'], g = _bn_bwd(
            dout, rec['out'], relu, rec['y2'], rec['m2'], rec['r2'], P[p + '.bn2.weight'],
            c2.rows_out, c2.Co, st, want_g=not has_ds, out_hi=rec['out_hi'], ws=ws_out, **kw)
        ws_out = None
        dy2 = op(dy2r, dy2p)
        if has_ds:
            cd = rec['cd']
            dydr, dydp, G[p + '.downsample.1.weight'], G[p + '.downsample.1.bias'], _ = _bn_bwd(
                dout, rec['out'], relu, rec['yd'], rec['md'], rec['rd'], P[p + '.downsample.1.weight'],
                cd.rows_out, cd.Co, st, out_hi=rec['out_hi'], **kw)
            dyd = op(dydr, dydp)
        del dout
        wq.add(G, p + '.conv2.weight', c2, rec['a1_op'], dy2)
        # conv2 is always stride 1: its dgrad also reduces bn1's backward sums in the epilogue (FUSE_BN_REDUCE)
        ws1 = None
        if tc and c2.fuse_bnred:
            da1, ws1 = c2.dgrad_bnred(dy2, st, rec['a1_op'][0], rec['y1'], rec['m1'], rec['r1'])
            wq.flush()
        else:
            da1 = c2.dgrad(dy2, st)
            wq.flush()
        del dy2, dy2r, dy2p
        dy1r, dy1p, G[p + '.bn1.weight'], G[p + '.bn1.bias'], _ = _bn_bwd(
            da1, rec['a1'], True, rec['y1'], rec['m1'], rec['r1'], P[p + '.bn1.weight'],
            c1.rows_out, c1.Co, st, out_hi=(rec['a1_op'][0] if tc else None), ws=ws1, **kw)
        dy1 = op(dy1r, dy1p)
        del da1
        wq.add(G, p + '.conv1.weight', c1, rec['xin_op'], dy1)
        if has_ds:
            dx = c1.dgrad(dy1, st)
            cd.dgrad(dyd, st, dx=dx)
            wq.add(G, p + '.downsample.0.weight', cd, rec['xin_op'], dyd)
            wq.flush()
            del dyd, dydr, dydp
        elif tc and c1.fuse_bnred and bi > 0 and blocks[bi - 1].get('out_hi') is not None \
                and blocks[bi - 1]['spec']['block'] == 'basic':
            # dx = g + dgrad is the previous block's output gradient: reduce that block's bn2 sums here
            prev = blocks[bi - 1]
            dx, ws_out = c1.dgrad_bnred(dy1, st, prev['out_hi'] if prev['spec']['final_relu'] else None,
                                        prev['y2'], prev['m2'], prev['r2'], dx=g)
            wq.flush()
        else:
            dx = c1.dgrad(dy1, st, dx=g)          # dx = g + dgrad
            wq.flush()
        del dy1, dy1r, dy1p
        dout = dx
        rec.clear()
    NB, T, H, W, Ho, Wo = ctx['stem_dims']
    rows0 = NB * T * Ho * Wo
    if ctx.get('ypool') is not None:
        # pooled stem (stem_pool.cu): bn1's backward sums on the pooled grid, then conv1 is RECOMPUTED and its epilogue
        # turns the pooled gradient into the gradient planes of the conv1 grid -- the conv1 output was never stored
        dev = dout.device
        x2, ypool = ctx['x2'], ctx['ypool']
        rows_p = ypool.shape[0]
        ws = torch.empty(128, dtype=torch.float64, device=dev)
        G['bn1.weight'], G['bn1.bias'] = _empty((64,), ypool), _empty((64,), ypool)
        _timed('stem_tail_bwd')(L.stem_pool_bwd_reduce)(ptr(ypool), ptr(dout), ptr(ctx['idx']), ptr(ctx['m0']), ptr(ctx['r0']), ptr(ws),
                                                        ptr(G['bn1.weight']), ptr(G['bn1.bias']), rows_p, st)
        if ctx.get('frozen'):
            ws.zero_()                 # fixed (running) statistics: dy = gamma * rstd * g, no batch-statistics terms
        dw0 = torch.empty_like(P['conv1.weight'])
        if STEM_FUSE_WGRAD and L.stem_pool_supported(H, W) == 2:
            # conv1's wgrad MMAs run in the same kernel on the gradient tile in shared memory: no 5.4 GB gradient planes
            _timed('stem_bwd_wgrad')(L.stem_pool_bwd_wgrad)(ptr(x2[0]), ptr(x2[1]), ptr(ctx['wp']), ptr(dout), ptr(ctx['idx']),
                                                            ptr(ctx['m0']), ptr(ctx['r0']), ptr(P['bn1.weight']), ptr(ws), ptr(dw0),
                                                            NB, T, H, W, st)
            del dout
        else:
            dy0p = (torch.empty((rows0, 64), dtype=GRD, device=dev), torch.empty((rows0, 64), dtype=GRD, device=dev))
            _timed('stem_tail_bwd')(L.stem_pool_bwd)(ptr(x2[0]), ptr(x2[1]), ptr(ctx['wp']), ptr(dout), ptr(ctx['idx']), ptr(ctx['m0']),
                                                     ptr(ctx['r0']), ptr(P['bn1.weight']), ptr(ws), ptr(dy0p[0]), ptr(dy0p[1]),
                                                     NB, T, H, W, st)
            del dout
            _timed('stem_wgrad')(L.stem_conv_wgrad_s2d)(ptr(x2[0]), ptr(x2[1]), ptr(dy0p[0]), ptr(dy0p[1]), ptr(dw0), NB, T, H, W, st)
    else:
        if ctx.get('frozen'):
            raise NotImplementedError('backward through eval-mode BatchNorm needs the pooled stem (even frame sizes)')
        dw0 = _stem_backward_unpooled(L, st, ctx, dout, P, G)
        del dout
    G['conv1.weight'] = dw0
    wq.flush()
    if side is not None:
        main.wait_stream(side)
    return G


# ==================================================================================================
# head: pool/split + ConvGRU + predictor + score
# ==================================================================================================
HEAD_PARAM_NAMES = ['agg.cell_list.0.reset_gate.weight', 'agg.cell_list.0.reset_gate.bias',
                    'agg.cell_list.0.update_gate.weight', 'agg.cell_list.0.update_gate.bias',
                    'agg.cell_list.0.out_gate.weight', 'agg.cell_list.0.out_gate.bias',
                    'network_pred.0.weight', 'network_pred.0.bias',
                    'network_pred.2.weight', 'network_pred.2.bias']


def _gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, st, beta=0.0, a_off=0, b_off=0, c_off=0):
    es = 4
    lib().gemm_f32(ta, tb, M, N, K, 1.0, A.data_ptr() + a_off * es, lda, B.data_ptr() + b_off * es, ldb,
                   beta, C.data_ptr() + c_off * es, ldc, st)


class _Gru:
    """one ConvGRU cell with kernel_size 1, acting on [R, D] row matrices"""

    def __init__(self, P, D, st):
        self.Wr, self.br = P['agg.cell_list.0.reset_gate.weight'], P['agg.cell_list.0.reset_gate.bias']
        self.Wz, self.bz = P['agg.cell_list.0.update_gate.weight'], P['agg.cell_list.0.update_gate.bias']
        self.Wo, self.bo = P['agg.cell_list.0.out_gate.weight'], P['agg.cell_list.0.out_gate.bias']
        self.D, self.st = D, st

    def xproj(self, X, rows):
        """XP [rows, 3D] = X @ [Wx_z | Wx_r | Wx_o]^T   (x-part of the concatenated input, columns 0..D)"""
        D = self.D
        XP = _empty((rows, 3 * D), X)
        for gi, Wg in enumerate((self.Wz, self.Wr, self.Wo)):
            _gemm(0, 1, rows, D, D, X, D, Wg, 2 * D, XP, 3 * D, self.st, c_off=gi * D)
        return XP

    def step(self, XP, xp_row0, h, R, p, seed, offset):
        """XP rows [xp_row0, xp_row0+R) hold this step's x projections.  Returns (h_new, saved)."""
        L, D, st = lib(), self.D, self.st
        hzr = _empty((R, 2 * D), h)
        _gemm(0, 1, R, D, D, h, D, self.Wz, 2 * D, hzr, 2 * D, st, b_off=D)
        _gemm(0, 1, R, D, D, h, D, self.Wr, 2 * D, hzr, 2 * D, st, b_off=D, c_off=D)
        z, r, hr = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        base = XP.data_ptr() + xp_row0 * 3 * D * 4
        L.gru_gates_zr(base, base + D * 4, 3 * D, ptr(hzr), ptr(self.bz), ptr(self.br), ptr(h), ptr(z), ptr(r),
                       ptr(hr), R, D, st)
        ho = torch.empty_like(h)
        _gemm(0, 1, R, D, D, hr, D, self.Wo, 2 * D, ho, D, st, b_off=D)
        o, hn = torch.empty_like(h), torch.empty_like(h)
        keep = torch.empty_like(h) if p > 0 else None
        L.gru_out(base + 2 * D * 4, 3 * D, ptr(ho), ptr(self.bo), ptr(h), ptr(z), ptr(o), ptr(hn), ptr(keep),
                  float(p), seed, offset, R, D, st)
        return hn, dict(h=h, z=z, r=r, o=o, hr=hr, keep=keep)

    def step_bwd(self, sv, x, dhout, R, G):
        """x [R,D]: the step's input rows.  Accumulates weight grads into G; returns (dx, dh_prev)."""
        L, D, st = lib(), self.D, self.st
        h = sv['h']
        dpo, dzp, dh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        L.gru_bwd_out(ptr(dhout), ptr(sv['keep']), ptr(h), ptr(sv['z']), ptr(sv['o']), ptr(dpo), ptr(dzp),
                      ptr(dh), R, D, st)
        dhr = torch.empty_like(h)
        _gemm(0, 0, R, D, D, dpo, D, self.Wo, 2 * D, dhr, D, st, b_off=D)
        dzr = _empty((R, 2 * D), h)
        L.gru_bwd_zr(ptr(dhr), ptr(h), ptr(sv['r']), ptr(sv['z']), ptr(dzp), ptr(dzr), ptr(dh), R, D, st)
        # dh += dpre_z Wh_z + dpre_r Wh_r
        _gemm(0, 0, R, D, D, dzr, 2 * D, self.Wz, 2 * D, dh, D, st, beta=1.0, b_off=D)
        _gemm(0, 0, R, D, D, dzr, 2 * D, self.Wr, 2 * D, dh, D, st, beta=1.0, a_off=D, b_off=D)
        # dx = dpre_z Wx_z + dpre_r Wx_r + dpre_o Wx_o
        dx = torch.empty_like(h)
        _gemm(0, 0, R, D, D, dzr, 2 * D, self.Wz, 2 * D, dx, D, st)
        _gemm(0, 0, R, D, D, dzr, 2 * D, self.Wr, 2 * D, dx, D, st, beta=1.0, a_off=D)
        _gemm(0, 0, R, D, D, dpo, D, self.Wo, 2 * D, dx, D, st, beta=1.0)
        # weight grads: dW[:, :D] += dpre^T x ; dW[:, D:] += dpre^T h (hr for the out gate)
        for name, A, lda, aoff, hin in (('update_gate', dzr, 2 * D, 0, h), ('reset_gate', dzr, 2 * D, D, h),
                                         ('out_gate', dpo, D, 0, sv['hr'])):
            dW = G['agg.cell_list.0.%s.weight' % name]
            _gemm(1, 0, D, D, R, A, lda, x, D, dW, 2 * D, st, beta=1.0, a_off=aoff)
            _gemm(1, 0, D, D, R, A, lda, hin, D, dW, 2 * D, st, beta=1.0, a_off=aoff, c_off=D)
        # biases: column sums of dpre (dzr is [R,2D]: z | r)
        L.colsum(ptr(dzr), R, 2 * D, ptr(G['_bzr']), 1, st)
        L.colsum(ptr(dpo), R, D, ptr(G['agg.cell_list.0.out_gate.bias']), 1, st)
        return dx, dh


# the recurrent head as ONE kernel per direction (head_chain.cu); False = the per-step GEMM + gate kernels above
HEAD_CHAIN = True


def _head_chain_ok(D):
    return HEAD_CHAIN and D == 256


def _head_forward_chain(z4, dims, B, N, pred_step, P, dropout_p, seed, need_ctx):
    """pool/split -> head_chain_fwd (GRU aggregation + prediction loop, one launch) -> score GEMM"""
    L = lib()
    st = _stream()
    To, Lh, Lw = dims
    S, D = Lh * Lw, z4.shape[1]
    NB, Tagg, R = B * N, N - pred_step, B * S
    T7, M = N - 1, B * pred_step * S
    finf_all = _empty((NB * S, D), z4)
    feat = _empty((NB * S, D), z4)
    L.pool_split_fwd(ptr(z4), ptr(finf_all), ptr(feat), NB, To, S, D, st)
    g = 'agg.cell_list.0.'
    Wz, Wr, Wo = P[g + 'update_gate.weight'], P[g + 'reset_gate.weight'], P[g + 'out_gate.weight']
    W0, W2 = P['network_pred.0.weight'], P['network_pred.2.weight']
    wt_zr, wt_o = _empty((2 * D, 2 * D), z4), _empty((2 * D, D), z4)
    w0t, w2t = _empty((D, D), z4), _empty((D, D), z4)
    L.head_chain_pack(ptr(Wz), ptr(Wr), ptr(Wo), ptr(W0), ptr(W2), ptr(wt_zr), ptr(wt_o), ptr(w0t), ptr(w2t), st)
    XH, XO = _empty((T7 * R, 2 * D), z4), _empty((T7 * R, 2 * D), z4)
    Z, Rg, O = _empty((T7 * R, D), z4), _empty((T7 * R, D), z4), _empty((T7 * R, D), z4)
    Keep = _empty((T7 * R, D), z4) if dropout_p > 0 else None
    U, Hp, Pp = _empty((pred_step * R, D), z4), _empty((pred_step * R, D), z4), _empty((pred_step * R, D), z4)
    pred_rows = _empty((M, D), z4)
    L.head_chain_fwd(ptr(feat), ptr(wt_zr), ptr(wt_o), ptr(w0t), ptr(w2t), ptr(P[g + 'update_gate.bias']),
                     ptr(P[g + 'reset_gate.bias']), ptr(P[g + 'out_gate.bias']), ptr(P['network_pred.0.bias']),
                     ptr(P['network_pred.2.bias']), B, N, S, pred_step, float(dropout_p), seed, ptr(XH), ptr(XO), ptr(Z), ptr(Rg),
                     ptr(O), ptr(Keep), ptr(U), ptr(Hp), ptr(Pp), ptr(pred_rows), st)
    finf_rows = _empty((M, D), z4)
    L.gather_rows(ptr(finf_all), ptr(finf_rows), M, D, pred_step * S, N * S, Tagg * S, st)
    score = _empty((M, M), z4)
    pp, fp = _split(pred_rows, st, f16=True), _split(finf_rows, st, f16=True)
    _timed('score_fwd')(L.score_matmul_tc)(M, M, D, ptr(pp[0]), ptr(pp[1]), ptr(fp[0]), ptr(fp[1]), 1, ptr(score), st)
    ctx = None
    if need_ctx:
        ctx = dict(chain=True, B=B, N=N, P=pred_step, S=S, D=D, To=To, R=R, M=M, finf_all=finf_all, XH=XH, XO=XO, Z=Z, Rg=Rg,
                   O=O, Keep=Keep, U=U, Hp=Hp, Pp=Pp, pred_rows=pred_rows, finf_rows=finf_rows)
    return score, ctx


def _wgrad_rows(x_rows, dy_rows, st):
    """dW [Co, Ci] = dy_rows^T . x_rows (reduction over the rows) on the tcgen05 wgrad kernel"""
    rows, Ci = x_rows.shape
    Co = dy_rows.shape[1]
    xp, dyp = _split(x_rows, st), _split(dy_rows, st)
    geom = ConvGeom(1, 1, 1, rows, Ci, 1, 1, rows, Co, 1, 1, 1, 1, 1, 1, 0, 0, 0)
    scratch = _empty((Co, Ci), x_rows)
    dw = _empty((Co, Ci), x_rows)
    lib().conv3d_wgrad_tc(geom, ptr(xp[0]), ptr(xp[1]), ptr(dyp[0]), ptr(dyp[1]), ptr(scratch), ptr(dw), st)
    return dw


def _head_backward_chain(ctx, dscore, P):
    L = lib()
    st = _stream()
    B, N, Pn, S, D, To, R, M = (ctx[k] for k in ('B', 'N', 'P', 'S', 'D', 'To', 'R', 'M'))
    NB, Tagg, T7 = B * N, N - Pn, N - 1
    dev = dscore.device
    g = 'agg.cell_list.0.'
    dpred_rows = _empty((M, D), dscore)
    dfinf_rows = _empty((M, D), dscore)
    if M % 64 == 0:
        dsp = _split(dscore, st)
        ftp = _split(ctx['finf_rows'].t().contiguous(), st)
        _timed('score_bwd')(L.gemm_nt_split_tc)(M, D, M, ptr(dsp[0]), ptr(dsp[1]), ptr(ftp[0]), ptr(ftp[1]), 0,
                                                ptr(dpred_rows), 0, st)
        pp = _split(ctx['pred_rows'], st)
        geom = ConvGeom(1, 1, 1, M, D, 1, 1, M, M, 1, 1, 1, 1, 1, 1, 0, 0, 0)
        scratch = _empty((M, D), dscore)
        _timed('score_bwd')(L.conv3d_wgrad_tc)(geom, ptr(pp[0]), ptr(pp[1]), ptr(dsp[0]), ptr(dsp[1]), ptr(scratch),
                                               ptr(dfinf_rows), st)
        del dsp
    else:                                          # K = M must be a multiple of 64 on the tensor-core GEMM: tiny test sizes
        _gemm(0, 0, M, D, M, dscore, M, ctx['finf_rows'], D, dpred_rows, D, st)
        _gemm(1, 0, M, D, M, dscore, M, ctx['pred_rows'], D, dfinf_rows, D, st)
    dfinf_all = torch.zeros((NB * S, D), dtype=torch.float32, device=dev)
    L.scatter_rows(ptr(dfinf_rows), ptr(dfinf_all), M, D, Pn * S, N * S, Tagg * S, 0, st)
    dfeat = torch.zeros((NB * S, D), dtype=torch.float32, device=dev)
    DZR, DO = _empty((T7 * R, 2 * D), dscore), _empty((T7 * R, D), dscore)
    DP, DU = _empty((Pn * R, D), dscore), _empty((Pn * R, D), dscore)
    L.head_chain_bwd(ptr(dpred_rows), ptr(P[g + 'update_gate.weight']), ptr(P[g + 'reset_gate.weight']),
                     ptr(P[g + 'out_gate.weight']), ptr(P['network_pred.0.weight']), ptr(P['network_pred.2.weight']),
                     B, N, S, Pn, ptr(ctx['XH']), ptr(ctx['Z']), ptr(ctx['Rg']), ptr(ctx['O']), ptr(ctx['Keep']), ptr(ctx['U']),
                     ptr(ctx['Pp']), ptr(dfeat), ptr(DZR), ptr(DO), ptr(DP), ptr(DU), st)
    G = {}
    # weight gradients: reductions over (steps x rows) on the tensor-core wgrad kernel; biases: column sums
    dWzr = _wgrad_rows(ctx['XH'], DZR, st)                                   # [2D (z | r), 2D (x | h)]
    G[g + 'update_gate.weight'] = dWzr[:D].reshape(P[g + 'update_gate.weight'].shape)
    G[g + 'reset_gate.weight'] = dWzr[D:].reshape(P[g + 'reset_gate.weight'].shape)
    G[g + 'out_gate.weight'] = _wgrad_rows(ctx['XO'], DO, st).reshape(P[g + 'out_gate.weight'].shape)
    G['network_pred.2.weight'] = _wgrad_rows(ctx['U'], DP, st).reshape(P['network_pred.2.weight'].shape)
    G['network_pred.0.weight'] = _wgrad_rows(ctx['Hp'], DU, st).reshape(P['network_pred.0.weight'].shape)
    bzr = torch.empty(2 * D, dtype=torch.float32, device=dev)
    L.colsum(ptr(DZR), T7 * R, 2 * D, ptr(bzr), 0, st)
    G[g + 'update_gate.bias'], G[g + 'reset_gate.bias'] = bzr[:D], bzr[D:]
    for name, src, rows in ((g + 'out_gate.bias', DO, T7 * R), ('network_pred.2.bias', DP, Pn * R), ('network_pred.0.bias', DU, Pn * R)):
        G[name] = torch.empty(D, dtype=torch.float32, device=dev)
        L.colsum(ptr(src), rows, D, ptr(G[name]), 0, st)
    dz4 = _empty((NB * To * S, D), dscore)
    L.pool_split_bwd(ptr(ctx['finf_all']), ptr(dfinf_all), ptr(dfeat), ptr(dz4), NB, To, S, D, st)
    return dz4, G


@_timed('head_fwd')
def head_forward(z4, dims, B, N, pred_step, P, dropout_p=0.0, seed=0, need_ctx=True):
    """z4: backbone output rows [B*N*To*S, D] (To temporal slices, S = L*L positions).
    Returns (score [M, M] with M = B*pred_step*S, ctx)."""
    if _head_chain_ok(z4.shape[1]):
        return _head_forward_chain(z4, dims, B, N, pred_step, P, dropout_p, seed, need_ctx)
    L = lib()
    st = _stream()
    To, Lh, Lw = dims
    S, D = Lh * Lw, z4.shape[1]              # feature size: 256 (r18 / r34) or 1024 (Bottleneck networks)
    NB = B * N
    Tagg = N - pred_step
    R = B * S
    finf_all = _empty((NB * S, D), z4)
    feat = _empty((NB * S, D), z4)
    L.pool_split_fwd(ptr(z4), ptr(finf_all), ptr(feat), NB, To, S, D, st)
    gru = _Gru(P, D, st)
    # aggregate: x_t rows (b, s) <- feat[(b*N + t)*S + s]
    X_all = _empty((Tagg * R, D), z4)
    for t in range(Tagg):
        L.gather_rows(ptr(feat), X_all.data_ptr() + t * R * D * 4, R, D, S, N * S, t * S, st)
    XP = gru.xproj(X_all, Tagg * R)
    h = torch.zeros((R, D), dtype=torch.float32, device=z4.device)
    steps = []
    off = 0
    for t in range(Tagg):
        h, sv = gru.step(XP, t * R, h, R, dropout_p, seed, off)
        off += R * D
        sv['x'] = None          # x_t = X_all[t]
        steps.append(sv)
    W0, b0 = P['network_pred.0.weight'], P['network_pred.0.bias']
    W2, b2 = P['network_pred.2.weight'], P['network_pred.2.bias']
    M = B * pred_step * S
    pred_rows = _empty((M, D), z4)
    psteps = []
    for i in range(pred_step):
        u = torch.empty_like(h)
        _gemm(0, 1, R, D, D, h, D, W0, D, u, D, st)
        L.bias_relu(ptr(u), ptr(b0), ptr(u), 1, R, D, st)
        pr = torch.empty_like(h)
        _gemm(0, 1, R, D, D, u, D, W2, D, pr, D, st)
        L.bias_relu(ptr(pr), ptr(b2), ptr(pr), 0, R, D, st)
        L.scatter_rows(ptr(pr), ptr(pred_rows), R, D, S, pred_step * S, i * S, 0, st)
        rec = dict(h=h, u=u, p=pr)
        if i < pred_step - 1:                     # the GRU step after the last prediction is dead work
            xr = torch.empty_like(h)
            L.bias_relu(ptr(pr), None, ptr(xr), 1, R, D, st)
            xp = gru.xproj(xr, R)
            h, sv = gru.step(xp, 0, h, R, dropout_p, seed, off)
            off += R * D
            sv['x'] = xr
            rec['gru'] = sv
        psteps.append(rec)
    finf_rows = _empty((M, D), z4)
    L.gather_rows(ptr(finf_all), ptr(finf_rows), M, D, pred_step * S, N * S, Tagg * S, st)
    score = _empty((M, M), z4)
    pp, fp = _split(pred_rows, st, f16=True), _split(finf_rows, st, f16=True)
    _timed('score_fwd')(L.gemm_nt_split_tc)(M, M, D, ptr(pp[0]), ptr(pp[1]), ptr(fp[0]), ptr(fp[1]), 1, ptr(score), 0, st)
    ctx = None
    if need_ctx:
        ctx = dict(B=B, N=N, P=pred_step, S=S, D=D, To=To, R=R, M=M, finf_all=finf_all, X_all=X_all,
                   steps=steps, psteps=psteps, pred_rows=pred_rows, finf_rows=finf_rows)
    return score, ctx


@_timed('head_bwd')
def head_backward(ctx, dscore, P):
    """returns (dz4 rows [B*N*To*S, D], grads dict for HEAD_PARAM_NAMES)"""
    if ctx.get('chain'):
        return _head_backward_chain(ctx, dscore, P)
    L = lib()
    st = _stream()
    B, N, Pn, S, D, To, R, M = (ctx[k] for k in ('B', 'N', 'P', 'S', 'D', 'To', 'R', 'M'))
    NB, Tagg = B * N, N - Pn
    dev = dscore.device
    G = {n: torch.zeros_like(P[n]) for n in HEAD_PARAM_NAMES}
    G['_bzr'] = torch.zeros(2 * D, dtype=torch.float32, device=dev)     # update | reset bias grads
    G['agg.cell_list.0.update_gate.bias'] = G['_bzr'][:D]
    G['agg.cell_list.0.reset_gate.bias'] = G['_bzr'][D:]
    gru = _Gru(P, D, st)
    W0, W2 = P['network_pred.0.weight'], P['network_pred.2.weight']
    dpred_rows = _empty((M, D), dscore)
    dfinf_rows = _empty((M, D), dscore)
    if M % 64 == 0:
        dsp = _split(dscore, st)
        # dpred = dS . finf      -> NT GEMM against finf^T (tiny transpose: data movement only)
        ftp = _split(ctx['finf_rows'].t().contiguous(), st)
        _timed('score_bwd')(L.gemm_nt_split_tc)(M, D, M, ptr(dsp[0]), ptr(dsp[1]), ptr(ftp[0]), ptr(ftp[1]), 0,
                                                ptr(dpred_rows), 0, st)
        # dfinf = dS^T . pred    -> the wgrad form (reduction over rows, both operands MN-major)
        pp = _split(ctx['pred_rows'], st)
        geom = ConvGeom(1, 1, 1, M, D, 1, 1, M, M, 1, 1, 1, 1, 1, 1, 0, 0, 0)
        scratch = _empty((M, D), dscore)
        _timed('score_bwd')(L.conv3d_wgrad_tc)(geom, ptr(pp[0]), ptr(pp[1]), ptr(dsp[0]), ptr(dsp[1]), ptr(scratch),
                                               ptr(dfinf_rows), st)
        del dsp
    else:
        _gemm(0, 0, M, D, M, dscore, M, ctx['finf_rows'], D, dpred_rows, D, st)
        _gemm(1, 0, M, D, M, dscore, M, ctx['pred_rows'], D, dfinf_rows, D, st)
    dfinf_all = torch.zeros((NB * S, D), dtype=torch.float32, device=dev)
    L.scatter_rows(ptr(dfinf_rows), ptr(dfinf_all), M, D, Pn * S, N * S, Tagg * S, 0, st)
    dfeat = torch.zeros((NB * S, D), dtype=torch.float32, device=dev)
    dh = None
    for i in reversed(range(Pn)):
        rec = ctx['psteps'][i]
        dp = _empty((R, D), dscore)
        L.gather_rows(ptr(dpred_rows), ptr(dp), R, D, S, Pn * S, i * S, st)
        if 'gru' in rec:
            sv = rec['gru']
            dxr, dh_prev = gru.step_bwd(sv, sv['x'], dh, R, G)
            L.relu_bwd(ptr(rec['p']), ptr(dxr), ptr(dp), 1, R * D, st)       # dp += dxr * (p > 0)
            dh = dh_prev
        # p = u W2^T + b2 ; u = relu(h W0^T + b0)
        _gemm(1, 0, D, D, R, dp, D, rec['u'], D, G['network_pred.2.weight'], D, st, beta=1.0)
        L.colsum(ptr(dp), R, D, ptr(G['network_pred.2.bias']), 1, st)
        du = _empty((R, D), dscore)
        _gemm(0, 0, R, D, D, dp, D, W2, D, du, D, st)
        L.relu_bwd(ptr(rec['u']), ptr(du), ptr(du), 0, R * D, st)
        _gemm(1, 0, D, D, R, du, D, rec['h'], D, G['network_pred.0.weight'], D, st, beta=1.0)
        L.colsum(ptr(du), R, D, ptr(G['network_pred.0.bias']), 1, st)
        if dh is None:
            dh = _empty((R, D), dscore)
            _gemm(0, 0, R, D, D, du, D, W0, D, dh, D, st)
        else:
            _gemm(0, 0, R, D, D, du, D, W0, D, dh, D, st, beta=1.0)
    X_all = ctx['X_all']
    for t in reversed(range(Tagg)):
        x_t = X_all[t * R:(t + 1) * R]
        dx, dh = gru.step_bwd(ctx['steps'][t], x_t, dh, R, G)
        L.scatter_rows(ptr(dx), ptr(dfeat), R, D, S, N * S, t * S, 0, st)
    dz4 = _empty((NB * To * S, D), dscore)
    L.pool_split_bwd(ptr(ctx['finf_all']), ptr(dfinf_all), ptr(dfeat), ptr(dz4), NB, To, S, D, st)
    del G['_bzr']
    return dz4, G


# ==================================================================================================
# general ConvGRU sequence (kernel_size 1, one layer) and the LC classifier head
#   ConvGRU.forward         /root/reference/backbone/convrnn.py:62-88
#   LC.forward (after the backbone)   /root/reference/eval/model_3d_lc.py:53-65
# ==================================================================================================
def gru_sequence_forward(X_all, h0, P, R, T, dropout_p, seed):
    """X_all rows [T*R, D] (step-major); h0 [R, D] or None.  Returns (H_all [T*R, D] post-dropout states, steps)."""
    st = _stream()
    D = X_all.shape[1]
    gru = _Gru(P, D, st)
    XP = gru.xproj(X_all, T * R)
    h = h0 if h0 is not None else torch.zeros((R, D), dtype=torch.float32, device=X_all.device)
    H_all = _empty((T * R, D), X_all)
    steps = []
    for t in range(T):
        h, sv = gru.step(XP, t * R, h, R, dropout_p, seed, t * R * D)
        H_all[t * R:(t + 1) * R].copy_(h)
        steps.append(sv)
    return H_all, steps


def gru_sequence_backward(steps, X_all, dH_all, dh_last, P, R, T, G):
    """dH_all [T*R, D] or None: gradient w.r.t. every step's output; dh_last [R, D] or None: extra gradient on the
    final state.  Accumulates weight grads into G.  Returns (dX_all [T*R, D], dh0)."""
    st = _stream()
    D = X_all.shape[1]
    gru = _Gru(P, D, st)
    L = lib()
    dX_all = _empty((T * R, D), X_all)
    dh = None
    if dh_last is not None:
        dh = torch.empty_like(dh_last)
        dh.copy_(dh_last)
    for t in reversed(range(T)):
        if dH_all is not None:
            cur_ptr = dH_all.data_ptr() + t * R * D * 4
            if dh is None:
                dh = _empty((R, D), X_all)
                L.gather_rows(cur_ptr, ptr(dh), R, D, R, R, 0, st)
            else:
                L.scatter_rows(cur_ptr, ptr(dh), R, D, R, R, 0, 1, st)           # dh += dH_all[t]
        if dh is None:
            dh = torch.zeros((R, D), dtype=torch.float32, device=X_all.device)
        dx, dh = gru.step_bwd(steps[t], X_all[t * R:(t + 1) * R], dh, R, G)
        dX_all[t * R:(t + 1) * R].copy_(dx)
    return dX_all, dh


def _new_head_grads(P, dev):
    D = P['agg.cell_list.0.update_gate.bias'].numel()
    G = {n: torch.zeros_like(P[n]) for n in HEAD_PARAM_NAMES[:6]}
    G['_bzr'] = torch.zeros(2 * D, dtype=torch.float32, device=dev)
    G['agg.cell_list.0.update_gate.bias'] = G['_bzr'][:D]
    G['agg.cell_list.0.reset_gate.bias'] = G['_bzr'][D:]
    return G


LC_PARAM_NAMES = HEAD_PARAM_NAMES[:6] + ['final_bn.weight', 'final_bn.bias', 'final_fc.1.weight', 'final_fc.1.bias']


def lc_head_forward(z4, dims, B, N, P, final_bn_state, training, gru_p, fc_p, seed, need_ctx=True):
    """z4: backbone rows [B*N*To*S, D].  Returns (output [B, num_class], context [B, D], ctx)."""
    L = lib()
    st = _stream()
    To, Lh, Lw = dims
    S, D = Lh * Lw, z4.shape[1]
    NB, R = B * N, B * S
    feat = _empty((NB * S, D), z4)
    L.relu_pool_fwd(ptr(z4), ptr(feat), NB, To, S * D, st)                   # ReLU, then temporal mean
    X_all = _empty((N * R, D), z4)
    for t in range(N):
        L.gather_rows(ptr(feat), X_all.data_ptr() + t * R * D * 4, R, D, S, N * S, t * S, st)
    H_all, steps = gru_sequence_forward(X_all, None, P, R, N, gru_p if training else 0.0, seed)
    h_last = H_all[(N - 1) * R:]
    vec = _empty((B, D), z4)
    dummy = _empty((B, D), z4)
    L.pool_split_fwd(ptr(h_last), ptr(vec), ptr(dummy), B, S, 1, D, st)       # spatial mean over the S positions
    if training:
        mean, rstd = _bn_stats(vec, B, D, st)
        L.bn_running_update(ptr(mean), ptr(rstd), B, BN_EPS, BN_MOMENTUM, ptr(final_bn_state[0]), ptr(final_bn_state[1]), D, st)
    else:
        mean = final_bn_state[0]
        rstd = torch.empty_like(mean)
        L.bn_rstd_from_var(ptr(final_bn_state[1]), BN_EPS, ptr(rstd), D, st)
    context, _ = _bn_apply(vec, mean, rstd, P['final_bn.weight'], P['final_bn.bias'], False, B, D, st)
    keep = None
    cin = context
    if training and fc_p > 0:
        cin, keep = torch.empty_like(context), torch.empty_like(context)
        L.dropout_fwd(ptr(context), ptr(cin), ptr(keep), float(fc_p), seed ^ 0x5DEECE66D, 0, B * D, st)
    W, bias = P['final_fc.1.weight'], P['final_fc.1.bias']
    nc = W.shape[0]
    out = _empty((B, nc), z4)
    _gemm(0, 1, B, nc, D, cin, D, W, D, out, nc, st)
    L.bias_relu(ptr(out), ptr(bias), ptr(out), 0, B, nc, st)
    ctx = None
    if need_ctx:
        ctx = dict(B=B, N=N, S=S, D=D, To=To, R=R, z4=z4, X_all=X_all, steps=steps, vec=vec, mean=mean, rstd=rstd,
                   keep=keep, cin=cin, nc=nc, frozen=not training)
    return out, context, ctx


def lc_head_backward(ctx, dout, dcontext, P):
    L = lib()
    st = _stream()
    B, N, S, D, To, R, nc = (ctx[k] for k in ('B', 'N', 'S', 'D', 'To', 'R', 'nc'))
    NB = B * N
    dev = dout.device
    G = _new_head_grads(P, dev)
    W = P['final_fc.1.weight']
    G['final_fc.1.weight'] = _empty((nc, D), dout)
    _gemm(1, 0, nc, D, B, dout, nc, ctx['cin'], D, G['final_fc.1.weight'], D, st)
    G['final_fc.1.bias'] = _empty((nc,), dout)
    L.colsum(ptr(dout), B, nc, ptr(G['final_fc.1.bias']), 0, st)
    dc = _empty((B, D), dout)
    _gemm(0, 0, B, D, nc, dout, nc, W, D, dc, D, st)
    if ctx['keep'] is not None:
        L.mul(ptr(dc), ptr(ctx['keep']), ptr(dc), B * D, st)
    if dcontext is not None:
        L.scatter_rows(ptr(dcontext.contiguous()), ptr(dc), B, D, B, B, 0, 1, st)      # dc += dcontext
    dvec, _, G['final_bn.weight'], G['final_bn.bias'], _ = _bn_bwd(dc, None, False, ctx['vec'], ctx['mean'], ctx['rstd'],
                                                                 P['final_bn.weight'], B, D, st, frozen=ctx.get('frozen', False))
    dh_last = _empty((R, D), dout)
    L.pool_split_bwd(ptr(ctx['vec']), ptr(dvec), None, ptr(dh_last), B, S, 1, D, st)
    dX_all, _ = gru_sequence_backward(ctx['steps'], ctx['X_all'], None, dh_last, P, R, N, G)
    dfeat = _empty((NB * S, D), dout)
    for t in range(N):
        L.scatter_rows(dX_all.data_ptr() + t * R * D * 4, ptr(dfeat), R, D, S, N * S, t * S, 0, st)
    dz4 = torch.empty_like(ctx['z4'])
    L.relu_pool_bwd(ptr(ctx['z4']), ptr(dfeat), ptr(dz4), NB, To, S * D, st)
    del G['_bzr']
    return dz4, G


# ==================================================================================================
# NCE mask / loss
# ==================================================================================================
def nce_mask(B, P, SQ, device):
    m = torch.empty((B, P, SQ, B, P, SQ), dtype=torch.int8, device=device)
    lib().nce_mask_fill(ptr(m), B, P, SQ, _stream())
    return m


def nce_ce_forward(score2d):
    rows, M = score2d.shape
    lse = _empty((rows,), score2d)
    out = _empty((4,), score2d)
    lib().nce_ce_fwd(ptr(score2d), rows, M, ptr(lse), ptr(out), _stream())
    return out, lse


def nce_ce_backward(score2d, lse, gscale):
    d = torch.empty_like(score2d)
    lib().nce_ce_bwd(ptr(score2d), ptr(lse), ptr(gscale), ptr(d), score2d.shape[0], score2d.shape[1], _stream())
    return d
