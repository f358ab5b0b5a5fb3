"""Worker of tests/test_multigpu.py, launched with torchrun (one process per GPU, NCCL over NVLink):

    each rank: same seeded DPC_RNN, its shard of a fixed global batch, one training step through FlatTrainer
               (backward into the flat gradient buffer, ONE all-reduce, Adam with the 1/world average folded in)
    rank 0   : (a) gathers every rank's parameters -> must be bit-identical across ranks;
               (b) replays the step in ONE process (both shards through the same model, gradients summed, Adam with the
                   same 1/world scale) -> parameters must agree with the distributed result.
Reference semantics: nn.DataParallel(model) + loss.backward() + optimizer.step(), /root/reference/dpc/main.py:65,229-231
(per-replica BatchNorm statistics and score matrix, gradient of the global-mean loss)."""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def build(dev, seed):
    import dpc_b200
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = dpc_b200.DPC_RNN(64, num_seq=8, seq_len=5, pred_step=3, network='resnet18')
    return m.to(dev).eval()                      # eval: dropout off (deterministic); BatchNorm still batch statistics


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import dpc_b200
    # rank-dependent seed on purpose: FlatTrainer must broadcast rank 0's parameters
    m = build(dev, 100 + rank)
    tr = dpc_b200.FlatTrainer(m, lr=1e-3, weight_decay=1e-5)
    p0 = tr.flat_p.clone()
    per = 2
    g = torch.Generator().manual_seed(7)
    block = torch.randn(per * world, 8, 3, 5, 64, 64, generator=g)
    crit = dpc_b200.NCECriterion()
    tr.zero_grad()
    loss = crit(m(block[rank * per:(rank + 1) * per].to(dev))[0])
    loss.backward()
    tr.step()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(tr.flat_p) for _ in range(world)]
    dist.all_gather(gathered, tr.flat_p)
    g_dist = tr.flat_g.clone()                               # the all-reduced (summed) gradient
    init = [torch.empty_like(p0) for _ in range(world)]
    dist.all_gather(init, p0)
    res = None
    if rank == 0:
        same_init = all(torch.equal(init[0], t) for t in init[1:])
        same_after = all(torch.equal(gathered[0], t) for t in gathered[1:])
        # single-process replay: same initial parameters (rank 0's), both shards, summed gradients, 1/world in Adam
        m2 = build(dev, 100)
        tr2 = dpc_b200.FlatTrainer(m2, lr=1e-3, weight_decay=1e-5, distributed=False)       # no collectives
        assert torch.equal(tr2.flat_p, p0)
        tr2.zero_grad()
        for r in range(world):
            crit(m2(block[r * per:(r + 1) * per].to(dev))[0]).backward()
        tr2.world = world                                    # Adam scale 1/world; allreduce() stays local ...
        tr2.allreduce = lambda: None                         # ... (the two shards' gradients were summed in place)
        tr2.step()
        torch.cuda.synchronize()
        upd_d = (gathered[0] - p0).double()
        upd_s = (tr2.flat_p - p0).double()
        rel = float((upd_d - upd_s).norm() / upd_s.norm())
        grel = float((g_dist.double() - tr2.flat_g.double()).norm() / tr2.flat_g.double().norm())
        res = dict(world=world, direct_nccl=tr.comm is not None, same_init=bool(same_init), same_after=bool(same_after), grad_rel_l2=grel, update_rel_l2=rel,
                   update_norm=float(upd_s.norm()), loss=float(loss))
        print('MPRESULT ' + json.dumps(res), flush=True)
    dist.barrier()
    tr.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
