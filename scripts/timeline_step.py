"""per-kernel-family timeline of ONE training step with the backward's chain / side streams intact: which weight-gradient
kernels (side stream) actually run under which chain kernels.   python scripts/timeline_step.py [B]"""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpc_b200
from dpc_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = dpc_b200.DPC_RNN(128, num_seq=8, seq_len=5, pred_step=3, network='resnet18').cuda().train()
crit = dpc_b200.NCECriterion()
tr = dpc_b200.FlatTrainer(m)
x = torch.randn(B, 8, 3, 5, 128, 128, device='cuda')


def step():
    tr.zero_grad()
    loss = crit(m(x)[0])
    loss.backward()
    tr.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print('untraced: %.2f ms/step' % (e0.elapsed_time(e1) / 5))
timer = engine.EventTimer(keep_overlap=True)
engine.set_timer(timer)
origin = torch.cuda.Event(enable_timing=True)
origin.record()
step()
end = torch.cuda.Event(enable_timing=True)
end.record()
engine.set_timer(None)
torch.cuda.synchronize()
print('traced step: %.2f ms' % origin.elapsed_time(end))
ids = {}
rows = []
for (tag, a, b, d), st in zip(timer.records, timer.streams):
    sid = ids.setdefault(st, len(ids))
    rows.append((origin.elapsed_time(a), origin.elapsed_time(b), sid, tag, d))
rows.sort()
side = [r for r in rows if r[3] == 'conv_wgrad' or r[3] == 'stem_wgrad']
chain = [r for r in rows if r[2] != (side[0][2] if side else -1)]
print('streams:', ids)
tot = {}
for s0, s1, sid, tag, d in rows:
    tot[(sid, tag)] = tot.get((sid, tag), 0.0) + (s1 - s0)
for k, v in sorted(tot.items()):
    print('stream %d %-16s %.2f ms' % (k[0], k[1], v))
# how much of each wgrad interval overlaps chain kernels, by chain tag
ov = {}
for s0, s1, sid, tag, d in side:
    for c0, c1, cid, ctag, cd in chain:
        lo, hi = max(s0, c0), min(s1, c1)
        if hi > lo:
            ov[ctag] = ov.get(ctag, 0.0) + hi - lo
print('wgrad time on the side stream: %.2f ms; overlapped with chain kernels by tag:' % sum(r[1] - r[0] for r in side))
for k, v in sorted(ov.items(), key=lambda kv: -kv[1]):
    print('   %-16s %.2f ms' % (k, v))
bw = [r for r in rows if r[3] in ('conv_dgrad', 'bn_bwd', 'conv_wgrad', 'stem_bwd_wgrad', 'stem_tail_bwd', 'head_bwd', 'score_bwd')]
if bw:
    print('backward window: %.2f .. %.2f ms' % (min(r[0] for r in bw), max(r[1] for r in bw)))
if '-v' in sys.argv:
    for s0, s1, sid, tag, d in rows:
        print('%8.3f %8.3f  s%d %-16s %s' % (s0, s1, sid, tag, d if d else ''))
