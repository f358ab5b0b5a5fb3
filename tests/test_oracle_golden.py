"""CPU: the oracle restatement against the golden vectors produced by the live reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no tests)."""
import math

import pytest
import torch

from oracle import dpc_oracle as O
from tests.util import CASES, ORACLE_ONLY_CASES, load_fixture, make_block, rel_err, check_sample

TOL = 2e-5          # fp32 CPU vs fp32 CPU, same ATen ops, different composition (functional vs modules)


@pytest.mark.parametrize('case', CASES + ORACLE_ONLY_CASES)
def test_forward_matches_reference(case):
    fx = load_fixture(case)
    sd = O.synthetic_state_dict(fx['network'], fx['seed_w'])
    assert list(sd.keys()) == fx['param_keys']
    block = make_block(fx)
    taps = {}
    score, mask = O.dpc_forward(block, sd, fx['network'], fx['pred_step'], taps=taps)
    assert torch.equal(mask, fx['mask'])                         # bit-exact int8, and contiguous
    assert mask.is_contiguous() and mask.dtype == torch.int8
    e, l2 = rel_err(score, fx['score'])
    assert e < TOL and l2 < TOL, (e, l2)
    for k, s in fx['taps'].items():
        check_sample(taps[k], s, TOL, k)
    loss, sflat, target = O.nce_loss(score, mask)
    assert torch.equal(target, fx['target'])
    M = sflat.shape[0]
    assert torch.equal(target, torch.arange(M))                  # positives are the diagonal
    assert abs(float(loss) - fx['loss']) < 1e-5 * max(1.0, abs(fx['loss']))
    tk = [float(x) for x in O.topk_accuracy(sflat, target)]
    assert tk == pytest.approx(fx['topk'], abs=1e-6)


@pytest.mark.parametrize('case', ['r18_img64_b2', 'r34_img64_b3', 'r50_img64_b2'])
def test_grads_match_reference(case):
    fx = load_fixture(case)
    sd = O.synthetic_state_dict(fx['network'], fx['seed_w'])
    block = make_block(fx)
    loss, score, grads = O.train_step_grads(block, sd, fx['network'], fx['pred_step'])
    assert abs(float(loss) - fx['loss']) < 1e-5 * max(1.0, abs(fx['loss']))
    n = 0
    for k, s in fx['grads'].items():
        if s is None:
            assert grads[k] is None
            continue
        check_sample(grads[k], s, 5e-4, k)       # bwd sums in a different order; BN cancels heavily
        n += 1
    assert n >= 76 - 6                           # R18: 76 tensors, GRU registered twice (trap 6)


def test_mask_closed_form_values():
    m = O.closed_form_mask(3, 3, 2)
    assert m.shape == (3, 3, 4, 3, 3, 4)
    assert int((m == 1).sum()) == 3 * 3 * 4
    assert int((m == -1).sum()) == 3 * 4 * 3 * 2
    assert int((m == -3).sum()) == 3 * 3 * 4 * 3 * 3
    assert int((m == 0).sum()) == m.numel() - 36 - 72 - 324


@pytest.mark.parametrize('net', ['resnet18', 'resnet34', 'resnet50'])
def test_reference_init_restatement(net):
    """a16: kaiming-normal(fan_out) convs are bit-reproducible; the orthogonal GRU/pred weights go
    through LAPACK, so they are checked by property (W W^T = I) and, when the host matches, by value."""
    init = load_fixture('reference_init_seed0')[net]
    sd = O.reference_init_state_dict(net, 0)
    assert list(sd.keys()) == list(init.keys())
    for k, c in init.items():
        v = sd[k]
        assert tuple(v.shape) == tuple(c['shape'])
        if k.startswith('backbone.'):
            assert torch.equal(v.reshape(-1)[:8], c['head']), k
            assert float(v.double().sum()) == pytest.approx(c['sum'], rel=1e-9, abs=1e-9)
        elif k.endswith('weight'):
            w = v.reshape(v.shape[0], -1).double()
            assert float((w @ w.t() - torch.eye(w.shape[0], dtype=torch.float64)).abs().max()) < 1e-5
        else:
            assert float(v.abs().max()) == 0.0
    n_params = sum(v.numel() for k, v in sd.items() if not k.startswith('agg.ConvGRUCell_00'))
    assert n_params == {'resnet18': 14583104, 'resnet34': 32947776, 'resnet50': 31956032}[net]    # SURVEY §8 a1 (+ r50 probe)


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = {'a': torch.randn(7, 5), 'b': torch.randn(11)}
    q = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.Adam(q.values(), lr=1e-3, weight_decay=1e-5)
    st = {}
    for _ in range(3):
        g = {k: torch.randn_like(v) for k, v in p.items()}
        for k in q:
            q[k].grad = g[k].clone()
        opt.step()
        O.adam_step(p, g, st)
    for k in p:
        assert torch.allclose(p[k], q[k].detach(), rtol=1e-6, atol=1e-7)
