"""CPU: host-side logic -- module surface / state_dict compatibility, architecture tables, and the
data-parallel semantics (world_size-2 gloo) of SURVEY.md §8(e)."""
import io
import contextlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dpc_oracle as O


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize('net', ['resnet18', 'resnet34', 'resnet50'])
def test_module_surface_matches_reference_state_dict(net):
    import dpc_b200
    torch.manual_seed(0)
    m = _quiet(dpc_b200.DPC_RNN, 128, num_seq=8, seq_len=5, pred_step=3, network=net)
    sd = m.state_dict()
    shapes = O.param_shapes(net)
    assert list(sd.keys()) == list(shapes.keys())                    # incl. the duplicated GRU keys (trap 6)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert len(list(m.buffers())) == 0
    assert sd['agg.ConvGRUCell_00.reset_gate.weight'].data_ptr() == sd['agg.cell_list.0.reset_gate.weight'].data_ptr()
    # same init + RNG order as the reference (pinned through the oracle restatement and golden checks)
    ref = O.reference_init_state_dict(net, 0)
    for k in sd:
        if k.startswith('backbone.'):
            assert torch.equal(sd[k], ref[k]), k
    assert m.last_duration == 2 and m.last_size == 4
    fs = 1024 if net == 'resnet50' else 256                          # select_backbone.py:4-10
    assert m.param == {'feature_size': fs, 'num_layers': 1, 'hidden_size': fs}
    # loads a reference-style checkpoint
    m.load_state_dict(O.synthetic_state_dict(net, 3), strict=True)
    m.reset_mask()
    assert m.mask is None


def test_select_resnet_contract():
    import dpc_b200
    from dpc_b200.resnet_2d3d import neq_load_customized
    model, param = dpc_b200.select_resnet('resnet18', track_running_stats=False)
    assert param == {'feature_size': 256}
    assert model.out_dims(5, 128, 128) == (2, 4, 4) and model.out_dims(5, 224, 224) == (2, 7, 7)
    with pytest.raises(IOError):
        dpc_b200.select_resnet('vgg')
    m50, p50 = dpc_b200.select_resnet('resnet50', track_running_stats=False)
    assert p50 == {'feature_size': 1024} and m50.network == 'resnet50'
    part = {k: torch.zeros_like(v) for k, v in model.state_dict().items() if 'layer1' in k}
    part['not.a.key'] = torch.zeros(1)
    _quiet(neq_load_customized, model, part)
    assert float(model.state_dict()['layer1.0.conv1.weight'].abs().max()) == 0.0
    assert float(model.state_dict()['layer2.0.conv1.weight'].abs().max()) > 0.0


def test_arch_tables_agree_with_oracle():
    from dpc_b200.arch import backbone_spec
    for net in ('resnet18', 'resnet34', 'resnet50', 'resnet101'):
        a, b = backbone_spec(net), O.backbone_spec(net)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert all(x[k] == y[k] for k in y)


def test_shard_batch():
    from dpc_b200 import shard_batch
    assert [shard_batch(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dpc_b200 import shard_batch
    B, P, L, D = 4, 3, 2, 16
    g = torch.Generator().manual_seed(7)
    pred = torch.randn(B, P * L * L, D, generator=g)
    finf = torch.randn(B, P * L * L, D, generator=g)
    W = torch.randn(D, D, generator=g).requires_grad_(True)
    lo, hi = shard_batch(B, rank, world)
    # rank-local score matrix, mask and loss: no cross-rank negatives (main.py:180,212)
    b = hi - lo
    s = (pred[lo:hi].reshape(-1, D) @ W) @ finf[lo:hi].reshape(-1, D).t()
    loss, _, _ = O.nce_loss(s.view(b, P, L * L, b, P, L * L), O.closed_form_mask(b, P, L))
    loss.backward()
    flat = W.grad.reshape(-1).clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                      # the path's ONE collective
    flat /= world
    if rank == 0:
        out.put(flat)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_dataparallel_semantics():
    """average of rank-local gradients == gradient of the reference's DataParallel loss
    (per-replica score blocks, CE mean over all rows)."""
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, P, L, D = 4, 3, 2, 16
    g = torch.Generator().manual_seed(7)
    pred = torch.randn(B, P * L * L, D, generator=g)
    finf = torch.randn(B, P * L * L, D, generator=g)
    W = torch.randn(D, D, generator=g).requires_grad_(True)
    rows = []
    for r in range(world):                                           # DataParallel: replica r scores its own shard
        lo, hi = r * 2, r * 2 + 2
        rows.append((pred[lo:hi].reshape(-1, D) @ W) @ finf[lo:hi].reshape(-1, D).t())
    gathered = torch.cat(rows, 0)                                    # [B*P*SQ, B2*P*SQ], main.py:213
    target = torch.arange(gathered.shape[0]) % gathered.shape[1]
    torch.nn.functional.cross_entropy(gathered, target).backward()
    assert torch.allclose(got, W.grad.reshape(-1), rtol=1e-5, atol=1e-7)
