"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Test infrastructure.  Runs only in the build container (the reference does not exist on the GPU
box).  Usage:  python oracle/make_golden.py

The reference modules are imported from where they lie; the only plumbing is the CPU patch the
survey describes (SURVEY.md §3.4 trap 5): `torch.Tensor.cuda` -> identity for the hard-coded
`.cuda()` at model_3d.py:88 and convrnn.py:27.  No reference file is modified or copied.

Each fixture stores: config, seeds, the reference outputs that are small enough (score, mask,
loss, top-k, pooled features, pred), and strided samples + norms of the large ones (stage feature
maps, parameter gradients).  Parameters are NOT stored: they are regenerated from a seed with
oracle.dpc_oracle.synthetic_state_dict (portable, LAPACK-free).
"""
import os
import sys
import io
import contextlib

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = '/root/reference'
sys.path.insert(0, os.path.join(REF, 'backbone'))
sys.path.insert(0, os.path.join(REF, 'dpc'))

from oracle import dpc_oracle as O  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self          # CPU plumbing (trap 5)
with contextlib.redirect_stdout(io.StringIO()):
    import model_3d as ref_model_3d                      # noqa: E402  (the reference itself)


def sample(t, n=2048):
    """Deterministic strided sample of a tensor + its L2 norm and abs-max."""
    f = t.detach().reshape(-1).to(torch.float32)
    step = max(1, f.numel() // n)
    return dict(shape=tuple(t.shape), step=step, values=f[::step][:n].clone(),
                norm=float(f.double().norm()), absmax=float(f.abs().max()))


def build_reference(network, img, pred_step, sd):
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_model_3d.DPC_RNN(sample_size=img, num_seq=8, seq_len=5, network=network, pred_step=pred_step)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def make(network, img, B, seed_w, seed_x, pred_step=3, with_grads=True):
    sd = O.synthetic_state_dict(network, seed_w)
    m = build_reference(network, img, pred_step, sd)
    assert list(m.state_dict().keys()) == list(sd.keys()), 'state_dict key order differs'
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.eval()                                             # dropout off; BN still batch-stat (trap 1)
    g = torch.Generator().manual_seed(seed_x)
    block = torch.randn(B, 8, 3, 5, img, img, generator=g)

    # hooks on the reference backbone for the stage feature maps
    taps = {}
    hooks = []
    bb = m.backbone
    hooks.append(bb.conv1.register_forward_hook(lambda mod, i, o: taps.__setitem__('stem.conv', o.detach().clone())))
    hooks.append(bb.maxpool.register_forward_hook(lambda mod, i, o: taps.__setitem__('stem.out', o.detach().clone())))
    for li in range(1, 5):
        layer = getattr(bb, 'layer%d' % li)
        for bi, blk in enumerate(layer):
            name = 'layer%d.%d' % (li, bi)
            hooks.append(blk.register_forward_hook(
                lambda mod, i, o, name=name: taps.__setitem__(name, o.detach().clone())))
    score, mask = m(block)
    for h in hooks:
        h.remove()

    # driver-side loss restated from main.py:178-185,213-217 (.reshape instead of .view: trap 4)
    Bm, NP, SQ, B2, NS, _ = mask.size()
    target = (mask == 1)
    score_flat = score.reshape(Bm * NP * SQ, B2 * NS * SQ)
    target_flat = target.reshape(Bm * NP * SQ, B2 * NS * SQ).to(int).argmax(dim=1)
    loss = F.cross_entropy(score_flat, target_flat)
    maxk = 5
    _, pred = score_flat.topk(maxk, 1, True, True)
    correct = pred.t().eq(target_flat.view(1, -1).expand_as(pred.t()))
    topk = [float(correct[:k].contiguous().view(-1).float().sum(0) / target_flat.size(0)) for k in (1, 3, 5)]

    fx = dict(network=network, img=img, B=B, pred_step=pred_step, seed_w=seed_w, seed_x=seed_x,
              torch_version=torch.__version__,
              score=score.detach().clone(), mask=mask.detach().contiguous().clone(),
              target=target_flat.clone(), loss=float(loss), topk=topk,
              taps={k: sample(v) for k, v in taps.items()},
              param_keys=list(sd.keys()))
    # small full tensors
    fx['backbone_out'] = taps[[k for k in taps if k.startswith('layer4')][-1]].clone()
    if with_grads:
        m.zero_grad()
        loss.backward()
        grads = {}
        for k, p in m.named_parameters():
            grads[k] = sample(p.grad, 512) if p.grad is not None else None
        fx['grads'] = grads
    return fx


def reference_init_checks(network, seed=0):
    torch.manual_seed(seed)                              # main.py:50
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_model_3d.DPC_RNN(sample_size=128, num_seq=8, seq_len=5, network=network, pred_step=3)
    out = {}
    for k, v in m.state_dict().items():
        f = v.reshape(-1)
        out[k] = dict(shape=tuple(v.shape), sum=float(f.double().sum()), abssum=float(f.double().abs().sum()),
                      head=f[:8].clone())
    return out


def main():
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    cases = [
        ('r18_img64_b2', dict(network='resnet18', img=64, B=2, seed_w=11, seed_x=21)),
        ('r18_img128_b2', dict(network='resnet18', img=128, B=2, seed_w=12, seed_x=22)),
        ('r34_img64_b3', dict(network='resnet34', img=64, B=3, seed_w=13, seed_x=23)),
        ('r18_img96_b2_p2', dict(network='resnet18', img=96, B=2, seed_w=14, seed_x=24, pred_step=2)),
        # Bottleneck network: pins the oracle's r50 restatement (the product path does not build r50+ yet)
        ('r50_img64_b2', dict(network='resnet50', img=64, B=2, seed_w=15, seed_x=25)),
    ]
    for name, kw in cases:
        fx = make(**kw)
        path = os.path.join(outdir, name + '.pt')
        torch.save(fx, path)
        print(name, 'loss', fx['loss'], 'topk', fx['topk'], '%.1f KB' % (os.path.getsize(path) / 1e3))
    init = {net: reference_init_checks(net) for net in ('resnet18', 'resnet34', 'resnet50')}
    torch.save(init, os.path.join(outdir, 'reference_init_seed0.pt'))
    print('init checks saved')


if __name__ == '__main__' and '--lc' not in sys.argv:
    main()


def make_lc(network='resnet18', img=64, B=3, seed_w=51, seed_x=52, num_class=101):
    """golden vectors of the reference LC classifier (eval/model_3d_lc.py), eval and train (dropout p = 0) mode"""
    sys.path.insert(0, os.path.join(REF, 'eval'))
    with contextlib.redirect_stdout(io.StringIO()):
        import model_3d_lc as ref_lc
        m = ref_lc.LC(sample_size=img, num_seq=8, seq_len=5, network=network, dropout=0.0, num_class=num_class)
    sd = O.lc_synthetic_state_dict(network, seed_w, num_class)
    assert list(m.state_dict().keys()) == list(sd.keys()), 'LC state_dict key order differs'
    m.load_state_dict(sd, strict=True)
    m.agg.dropout_layer.p = 0.0
    g = torch.Generator().manual_seed(seed_x)
    block = torch.randn(B, 8, 3, 5, img, img, generator=g)
    fx = dict(network=network, img=img, B=B, seed_w=seed_w, seed_x=seed_x, num_class=num_class, keys=list(sd.keys()))
    m.eval()
    with torch.no_grad():
        out, ctxv = m(block)
    fx['eval_output'], fx['eval_context'] = out.clone(), ctxv.clone()
    m.train()
    out, ctxv = m(block)
    fx['train_output'], fx['train_context'] = out.detach().clone(), ctxv.detach().clone()
    target = torch.arange(B) % num_class
    loss = F.cross_entropy(out.view(B, num_class), target)
    loss.backward()
    fx['train_loss'] = float(loss.detach())
    fx['grads'] = {k: sample(p.grad, 512) for k, p in m.named_parameters()}
    new = m.state_dict()
    fx['new_stats'] = {k: new[k].clone() for k in new if 'running' in k and ('final_bn' in k or 'bn1.' in k[:13] or 'layer4.1' in k)}
    fx['num_batches_tracked'] = int(new['final_bn.num_batches_tracked'])
    return fx


if __name__ == '__main__' and '--lc' in sys.argv:
    for name, kw in (('lc_r18_img64_b3', dict()), ('lc_r50_img64_b4', dict(network='resnet50', B=4, seed_w=53, seed_x=54))):
        if '--only' in sys.argv and name not in sys.argv:
            continue
        fx = make_lc(**kw)
        path = os.path.join(ROOT, 'tests', 'golden', name + '.pt')
        torch.save(fx, path)
        print('lc fixture', name, fx['train_loss'], '%.1f KB' % (os.path.getsize(path) / 1e3))
