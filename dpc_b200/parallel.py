"""Data-parallel plumbing for the training step.

One process per GPU; parameters replicated; the batch dimension shards with no data-path
collective (score matrix, BN statistics, mask and loss stay per rank -- the reference's
nn.DataParallel semantics, /root/reference/dpc/main.py:65,180,212).  The only exchange is ONE
all-reduce (sum) of the flat fp32 gradient buffer over NCCL / NVLink, followed by a fused Adam
(torch.optim.Adam(lr, weight_decay) semantics, main.py:81) that folds the 1/world scaling in.

The all-reduce is issued by the library itself (`dpc_flat_allreduce`, csrc/comm.cu) on the compute stream, on a
communicator whose unique id travels through torch.distributed (the plumbing: rendezvous, the initial parameter
broadcast, barriers).  DPC_DIRECT_NCCL=0, a process sub-group, or a non-NCCL backend fall back to
`torch.distributed.all_reduce`.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from ._lib import lib, ptr


class FlatTrainer:
    """Owns flat views of a module's parameters and gradients.

    After construction every `p.data` / `p.grad` is a view into one flat fp32 buffer, so
    `loss.backward()` accumulates straight into the all-reduce buffer."""

    def __init__(self, module, lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, process_group=None,
                 distributed=True):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError('no trainable parameters')
        dev = params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FlatTrainer needs CUDA parameters')
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view(p.shape)
            p.grad = self.flat_g[off:off + k].view(p.shape)
            off += k
        self.params, self.n = params, n
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.step_count = 0
        self.group = process_group
        self.allreduce_events = None      # set to [] to collect a (start, end) CUDA event pair per all-reduce (bench.py)
        self.world = dist.get_world_size(process_group) if (distributed and dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            # replicas start identical whatever each rank's RNG did (nn.DataParallel re-broadcasts rank 0's parameters
            # every forward, main.py:65; one process per GPU needs it once)
            src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
            dist.broadcast(self.flat_p, src=src, group=process_group)
        self.comm = None
        if (self.world > 1 and process_group is None and dist.get_backend() == 'nccl'
                and os.environ.get('DPC_DIRECT_NCCL', '1') != '0'):
            self._init_comm(dev)

    def _init_comm(self, dev):
        """the library's own NCCL communicator: rank 0 draws the unique id, torch.distributed carries it"""
        rank = dist.get_rank()
        ident = ctypes.create_string_buffer(128)
        if rank == 0:
            lib().comm_unique_id(ctypes.addressof(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0)
        ident = ctypes.create_string_buffer(box[0], 128)
        comm = ctypes.c_void_p()
        with torch.cuda.device(dev):
            lib().comm_init(ctypes.addressof(ident), rank, self.world, ctypes.addressof(comm))
        self.comm = comm.value

    def close(self):
        if getattr(self, 'comm', None):
            lib().comm_destroy(self.comm)
            self.comm = None

    def zero_grad(self):
        self.flat_g.zero_()

    def allreduce(self):
        if self.world > 1:
            from . import engine
            tok = engine._TIMER.start('allreduce') if engine._TIMER is not None else None
            ev = None
            if self.allreduce_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if self.comm is not None:
                lib().flat_allreduce(self.comm, ptr(self.flat_g), self.n, torch.cuda.current_stream().cuda_stream)
            else:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            if ev is not None:
                ev[1].record()
                self.allreduce_events.append(ev)
            if tok is not None:
                engine._TIMER.stop(tok)

    def step(self):
        """all-reduce (sum) + Adam with the 1/world average folded into the update"""
        self.allreduce()
        self.step_count += 1
        lib().adam_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), self.n, self.lr,
                        self.betas[0], self.betas[1], self.eps, self.wd, self.step_count, 1.0 / self.world,
                        torch.cuda.current_stream().cuda_stream)


def shard_batch(global_batch, rank, world):
    """rows [lo, hi) of the global batch owned by `rank` (equal shards: drop_last semantics, main.py:313)"""
    if global_batch % world:
        raise ValueError('global batch %d is not divisible by world size %d' % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per
