"""halo-patch conv kernel vs the tap-per-box kernel: agreement (both descriptor modes) and timing"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpc_b200 import engine as E


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def check(NB, dims, baseoff):
    st = torch.cuda.current_stream().cuda_stream
    site = E.TcConvSite(NB, dims, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    g = torch.Generator(device='cuda').manual_seed(5)
    w = torch.randn(64, 64, 1, 3, 3, device='cuda', generator=g) * 0.05
    site.pack(w, st)
    x = torch.randn(site.rows_in, 64, device='cuda', generator=g)
    xp = E._split(x, st)
    dy = torch.randn(site.rows_out, 64, device='cuda', generator=g)
    dyp = E._split(dy, st)
    out = {}
    for mode in ('0', '1'):
        os.environ['DPC_TC_HALO'] = mode
        os.environ['DPC_HALO_BASEOFF'] = str(baseoff)
        y, mean, rstd = site.fwd_bn(xp, st)
        dx = site.dgrad(dyp, st)
        torch.cuda.synchronize()
        out[mode] = (y.clone(), torch.cat([mean, rstd]).clone(), dx.clone())
    print('dims %s NB %d baseoff %d: fwd rel %.3e  stats rel %.3e  dgrad rel %.3e' % (
        dims, NB, baseoff, rel(out['1'][0], out['0'][0]), rel(out['1'][1], out['0'][1]), rel(out['1'][2], out['0'][2])), flush=True)


def bench(NB, dims, iters=10):
    st = torch.cuda.current_stream().cuda_stream
    site = E.TcConvSite(NB, dims, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    w = torch.randn(64, 64, 1, 3, 3, device='cuda') * 0.05
    site.pack(w, st)
    x = torch.randn(site.rows_in, 64, device='cuda')
    xp = E._split(x, st)
    for mode in ('0', '1'):
        os.environ['DPC_TC_HALO'] = mode
        for name, fn in (('fwd_bn', lambda: site.fwd_bn(xp, st)), ('dgrad', lambda: site.dgrad(xp, st))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print('halo=%s %s dims %s NB %d: %.3f ms' % (mode, name, dims, NB, e0.elapsed_time(e1) / iters), flush=True)


if __name__ == '__main__':
    try:
        for bo in ((0,) if os.environ.get('PROBE_CHECK', '1') == '1' else ()):
            check(2, (3, 32, 32), bo)
            check(3, (2, 16, 16), bo)
            check(1, (2, 56, 56), bo)
    except Exception as e:                        # noqa: BLE001
        print('check failed:', e, flush=True)
    os.environ['DPC_HALO_BASEOFF'] = os.environ.get('PROBE_BASEOFF', '0')
    if os.environ.get('PROBE_BENCH', '1') == '1':
        bench(1024, (5, 32, 32))
        bench(220, (5, 56, 56))
