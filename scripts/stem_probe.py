"""stem conv1 forward: CUDA-core im2col + tcgen05 (stem_tc.cu) vs space-to-depth + TMA halo patch (stem_s2d.cu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpc_b200._lib import lib


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main(NB=1024, T=5, H=128, W=128):
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(NB, 3, T, H, W, device='cuda')
    w = torch.randn(64, 3, 1, 7, 7, device='cuda') * 0.1
    Ho, Wo = H // 2, W // 2
    y = torch.empty(NB * T * Ho * Wo, 64, device='cuda')
    y2 = torch.empty_like(y)
    ws = torch.empty(128, dtype=torch.float64, device='cuda')
    bf = dict(dtype=torch.bfloat16, device='cuda')
    x2h, x2l = torch.empty(NB, T, Ho, Wo, 16, **bf), torch.empty(NB, T, Ho, Wo, 16, **bf)
    wp = torch.empty(32768, **bf)
    t_old = timeit(lambda: L.stem_conv_fwd_tc(x.data_ptr(), w.data_ptr(), y.data_ptr(), ws.data_ptr(), NB, T, H, W, st))
    t_pack = timeit(lambda: L.stem_s2d_pack(x.data_ptr(), x2h.data_ptr(), x2l.data_ptr(), NB, T, H, W, st))
    t_new = timeit(lambda: L.stem_conv_fwd_s2d(x2h.data_ptr(), x2l.data_ptr(), w.data_ptr(), wp.data_ptr(), y2.data_ptr(),
                                                ws.data_ptr(), NB, T, H, W, st))
    torch.cuda.synchronize()
    err = float((y2 - y).abs().max() / y.abs().max())
    print('NB %d %dx%d: im2col kernel %.3f ms | s2d pack %.3f ms + conv %.3f ms | max rel diff %.2e' % (NB, H, W, t_old, t_pack, t_new, err), flush=True)


if __name__ == '__main__':
    main()
    main(NB=352, H=224, W=224)


def wgrad(NB=1024, T=5, H=128, W=128):
    from dpc_b200 import engine as E
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(NB, 3, T, H, W, device='cuda')
    Ho, Wo = H // 2, W // 2
    dy = torch.randn(NB * T * Ho * Wo, 64, device='cuda')
    dyp = E._split(dy, st)
    del dy
    bf = dict(dtype=torch.bfloat16, device='cuda')
    x2h, x2l = torch.empty(NB, T, Ho, Wo, 16, **bf), torch.empty(NB, T, Ho, Wo, 16, **bf)
    L.stem_s2d_pack(x.data_ptr(), x2h.data_ptr(), x2l.data_ptr(), NB, T, H, W, st)
    dw1, dw2 = torch.empty(64, 3, 1, 7, 7, device='cuda'), torch.empty(64, 3, 1, 7, 7, device='cuda')
    t_old = timeit(lambda: L.stem_conv_wgrad_tc(x.data_ptr(), dyp[0].data_ptr(), dyp[1].data_ptr(), dw1.data_ptr(), NB, T, H, W, st))
    t_new = timeit(lambda: L.stem_conv_wgrad_s2d(x2h.data_ptr(), x2l.data_ptr(), dyp[0].data_ptr(), dyp[1].data_ptr(), dw2.data_ptr(),
                                                  NB, T, H, W, st))
    torch.cuda.synchronize()
    print('wgrad NB %d %dx%d: im2col kernel %.3f ms | s2d %.3f ms | max rel diff %.2e' % (
        NB, H, W, t_old, t_new, float((dw2 - dw1).abs().max() / dw1.abs().max())), flush=True)


if __name__ == '__main__':
    wgrad()
