"""time the augmentation kernel at the BASELINE config-2 batch (128 clips x 40 frames of 200 x 150 -> 128^2):
   python scripts/bench_augment.py [B] [W] [H] [S]"""
import sys, os, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dpc_b200 import augmentation as D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
W = int(sys.argv[2]) if len(sys.argv) > 2 else 200
H = int(sys.argv[3]) if len(sys.argv) > 3 else 150
S = int(sys.argv[4]) if len(sys.argv) > 4 else 128
N, SL = 8, 5
frames = torch.randint(0, 256, (B, N * SL, H, W, 3), dtype=torch.uint8, device='cuda')
tr = D.k400_transform(S)
random.seed(0); np.random.seed(0)
t0 = time.time()
plans = [tr.plan(N * SL, W, H) for _ in range(B)]
t1 = time.time()
packed = tr.pack(plans)
t2 = time.time()
out = torch.empty(B, N, 3, SL, S, S, device='cuda')
ts = []
for i in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tr(frames, N, SL, plans=plans, out=out)
    b.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(a.elapsed_time(b))
ts.sort()
med = ts[len(ts) // 2]
# the kernel alone (tables already on the device)
from dpc_b200._lib import lib, ptr
tables, fpar, (Wo, Ho), K = packed
t_d, f_d = torch.from_numpy(tables).cuda(), torch.from_numpy(fpar).cuda()
mean, std = np.asarray(plans[0].normalize[0], np.float32), np.asarray(plans[0].normalize[1], np.float32)
ks = []
for i in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib().augment_clips(ptr(frames), ptr(t_d), ptr(f_d), mean.ctypes.data, std.ctypes.data, ptr(out), B, N * SL, H, W, Ho, Wo, K,
                        N, SL, torch.cuda.current_stream().cuda_stream)
    b.record()
    torch.cuda.synchronize()
    if i >= 2:
        ks.append(a.elapsed_time(b))
ks.sort()
print('kernel alone: median %.3f ms (%.1f GB/s), K = %d taps' % (ks[len(ks) // 2], (frames.numel() + out.numel() * 4) / ks[len(ks) // 2] / 1e6, K))
gb = (frames.numel() + out.numel() * 4) / 1e9
print('augment B=%d %dx%d -> %d^2: kernel+upload median %.3f ms (%.1f GB/s of frames in + block out, %.0f clips/s); '
      'host: draw %.1f ms, tables %.1f ms per batch' % (B, W, H, S, med, gb / med * 1e3, B / med * 1e3, (t1 - t0) * 1e3,
                                                        (t2 - t1) * 1e3))
