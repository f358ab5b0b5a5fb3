"""two training steps at BASELINE config 2 (the workload bench.py times), for ncu captures:
   ncu ... python scripts/profile_step.py [B]"""
import io, contextlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpc_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = dpc_b200.DPC_RNN(128, network='resnet18').cuda().train()
crit = dpc_b200.NCECriterion()
tr = dpc_b200.FlatTrainer(m)
x = torch.randn(B, 8, 3, 5, 128, 128, device='cuda')
for _ in range(2):
    tr.zero_grad()
    s, _ = m(x)
    loss = crit(s)
    loss.backward()
    tr.step()
torch.cuda.synchronize()
print('loss', float(loss))
