"""micro-benchmark of one conv site through the C ABI (fwd / dgrad / wgrad), CUDA-event timed"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpc_b200 import engine as E

def run(NB, dims, Ci, Co, k, s, p, iters=10):
    st = torch.cuda.current_stream().cuda_stream
    site = E.TcConvSite(NB, dims, Ci, Co, k, s, p)
    w = torch.randn(Co, Ci, *k, device='cuda') * 0.05
    site.pack(w, st)
    x = torch.randn(site.rows_in, Ci, device='cuda')
    xp = E._split(x, st)
    dy = torch.randn(site.rows_out, Co, device='cuda')
    dyp = E._split(dy, st)
    res = {}
    ybn = torch.randn(site.rows_in, Ci, device='cuda')
    mask = torch.randn(site.rows_in, Ci, device='cuda').to(torch.bfloat16)
    mean, rstd = torch.zeros(Ci, device='cuda'), torch.ones(Ci, device='cuda')
    for name, fn in (('fwd_bn', lambda: site.fwd_bn(xp, st)), ('dgrad', lambda: site.dgrad(dyp, st)),
                     ('dgrad_bnred', lambda: site.dgrad_bnred(dyp, st, mask, ybn, mean, rstd)),
                     ('wgrad', lambda: site.wgrad(xp, dyp, st))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2 * site.rows_out * Co * Ci * site.taps
        res[name] = (ms, fl / ms / 1e9)
    return res

if __name__ == '__main__':
    tag = os.environ.get('DPC_TC_RESIDENT', 'default')
    for name, args in (('layer1 64->64 1x3x3 32x32', (1024, (5, 32, 32), 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
                       ('layer2 128->128 1x3x3 16x16', (1024, (5, 16, 16), 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1))),
                       ('layer3 256->256 3x3x3 8x8', (1024, (3, 8, 8), 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1)))):
        r = run(*args)
        print('[resident=%s] %-30s ' % (tag, name) + '  '.join('%s %.3f ms (%.0f TF/s alg)' % (k, v[0], v[1]) for k, v in r.items()))
