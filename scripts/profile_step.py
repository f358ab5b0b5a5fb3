"""two training steps (the workload bench.py times), for ncu captures:
   ncu ... python scripts/profile_step.py [B] [network] [img]          default: BASELINE config 2 (128 resnet18 128)"""
import io, contextlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpc_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = sys.argv[2] if len(sys.argv) > 2 else 'resnet18'
img = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = dpc_b200.DPC_RNN(img, network=net).cuda().train()
crit = dpc_b200.NCECriterion()
tr = dpc_b200.FlatTrainer(m)
x = torch.randn(B, 8, 3, 5, img, img, device='cuda')
for _ in range(2):
    tr.zero_grad()
    s, _ = m(x)
    loss = crit(s)
    loss.backward()
    tr.step()
torch.cuda.synchronize()
print('loss', float(loss.detach()))
