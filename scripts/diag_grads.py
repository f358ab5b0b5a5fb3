"""diagnostic: per-parameter gradient error of the TC and SIMT paths vs the fp64 oracle; accumulator bias probe"""
import io, contextlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dpc_b200
from dpc_b200 import engine as E
from oracle import dpc_oracle as O
from tests.util import load_fixture, make_block

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
L = dpc_b200.lib()
st = torch.cuda.current_stream().cuda_stream

# ---- accumulator rounding probe: all-positive operands, long K
for K in (256, 4096, 65536):
    M = N = 128
    g = torch.Generator(device='cuda').manual_seed(0)
    A = torch.rand(M, K, device='cuda', generator=g) + 0.5
    B = torch.rand(N, K, device='cuda', generator=g) + 0.5
    ah, al = E._split(A, st); bh, bl = E._split(B, st)
    C = torch.empty(M, N, device='cuda')
    L.gemm_nt_split_tc(M, N, K, ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), 0, C.data_ptr(), 0, st)
    ref = A.double() @ B.double().t()
    ref16 = (ah.double() + al.double()) @ (bh.double() + bl.double()).t()
    c32 = A @ B.t()
    print('K=%6d  tc vs fp64: mean rel %.3e  (vs split-exact %.3e)   torch fp32 mean rel %.3e' % (
        K, float(((C.double() - ref) / ref).mean()), float(((C.double() - ref16) / ref16).mean()),
        float(((c32.double() - ref) / ref).mean())))

fx = load_fixture('r18_img64_b2')
sd = O.synthetic_state_dict(fx['network'], fx['seed_w'])
block = make_block(fx)
_, s64, g64 = O.train_step_grads(block.double(), {k: v.double() for k, v in sd.items()}, fx['network'], 3)
_, s32, g32 = O.train_step_grads(block, sd, fx['network'], 3)
res = {}
for tc in (False, True):
    E.USE_TC = tc
    with contextlib.redirect_stdout(io.StringIO()):
        m = dpc_b200.DPC_RNN(64, network='resnet18')
    m.load_state_dict(sd); m = m.cuda().eval()
    score, _ = m(block.cuda())
    dpc_b200.NCECriterion()(score).backward()
    print('tc=%s score rel err vs fp64 %.3e' % (tc, float((score.detach().cpu().double().reshape(-1) - s64.reshape(-1)).abs().max() / s64.abs().max())))
    res[tc] = {k: p.grad.detach().cpu().double() for k, p in m.named_parameters()}
print('%-42s %10s %10s %10s' % ('param', 'cpu32', 'simt', 'tc'))
for k in res[True]:
    ref = g64[k]
    n = lambda a: float((a - ref).norm() / ref.norm())
    print('%-42s %10.2e %10.2e %10.2e' % (k, n(g32[k].double()), n(res[False][k]), n(res[True][k])))
