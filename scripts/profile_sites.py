"""per-call-site CUDA-event timings of one training step (run on the GPU box)"""
import io, contextlib, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpc_b200
from dpc_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
with contextlib.redirect_stdout(io.StringIO()):
    m = dpc_b200.DPC_RNN(128, network='resnet18').cuda().train()
crit = dpc_b200.NCECriterion()
tr = dpc_b200.FlatTrainer(m)
x = torch.randn(B, 8, 3, 5, 128, 128, device='cuda')


def step():
    tr.zero_grad()
    s, _ = m(x)
    crit(s).backward()
    tr.step()


for _ in range(3):
    step()
# tag conv calls with their geometry
orig = {}
for cls in (E.TcConvSite,):
    for name in ('fwd_bn', 'fwd', 'dgrad', 'wgrad'):
        fn = getattr(cls, name)
        orig[(cls, name)] = fn

        def make(fn, name):
            def wrapped(self, *a, **k):
                g = self.geom
                tag = '%s Ci%d Co%d k%dx%dx%d s%d%d%d %dx%dx%d' % (name, g.Ci, g.Co, g.kT, g.kH, g.kW, g.sT, g.sH, g.sW, g.To, g.Ho, g.Wo)
                t = E._TIMER
                tok = t.start(tag)
                try:
                    return fn(self, *a, **k)
                finally:
                    t.stop(tok)
            return wrapped
        setattr(cls, name, make(fn, name))
timer = E.EventTimer()
E.set_timer(timer)
step()
E.set_timer(None)
tot = timer.totals()
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if ' Ci' in k:
        g = k.split()
        print('%-60s calls %2d  %8.3f ms' % (k, c, t))
