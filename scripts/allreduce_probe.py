"""torchrun --nproc-per-node N scripts/allreduce_probe.py : the 58 MB flat-gradient all-reduce through torch.distributed vs
through the library's own communicator (dpc_flat_allreduce), CUDA-event timed, same buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import dpc_b200
from dpc_b200._lib import lib, ptr

lin = torch.nn.Linear(3806, 3806).to(dev)           # ~14.5 M parameters, as the R18 DPC model
tr = dpc_b200.FlatTrainer(lin)
n = tr.n
buf = tr.flat_g
buf.fill_(1.0)


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


t_torch = timed(lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM))
buf.fill_(1.0)
st = torch.cuda.current_stream().cuda_stream
t_lib = timed(lambda: lib().flat_allreduce(tr.comm, ptr(buf), n, st)) if tr.comm is not None else None
buf.fill_(1.0)
lib().flat_allreduce(tr.comm, ptr(buf), n, st)
torch.cuda.synchronize()
ok = bool((buf == world).all())
if rank == 0:
    print('PROBE n=%d (%.1f MB) world=%d: torch.distributed %.3f ms, dpc_flat_allreduce %s ms, sum correct: %s'
          % (n, n * 4 / 1e6, world, t_torch, ('%.3f' % t_lib) if t_lib else None, ok), flush=True)
tr.close()
dist.destroy_process_group()
