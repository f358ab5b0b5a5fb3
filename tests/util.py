import os
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['r18_img64_b2', 'r18_img128_b2', 'r34_img64_b3', 'r18_img96_b2_p2', 'r50_img64_b2']
ORACLE_ONLY_CASES = []


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)


def make_block(fx):
    g = torch.Generator().manual_seed(fx['seed_x'])
    return torch.randn(fx['B'], 8, 3, 5, fx['img'], fx['img'], generator=g)


def rel_err(a, b):
    """max|a-b| / max|b|  (the north-star's relative fp32 tolerance measure) and rel-L2."""
    a = a.detach().double().reshape(-1).cpu()
    b = b.detach().double().reshape(-1).cpu()
    den = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / den, float((a - b).norm()) / max(float(b.norm()), 1e-30)


def check_sample(t, s, tol, what=''):
    """Compare tensor `t` with a fixture sample dict (see oracle/make_golden.py:sample)."""
    assert tuple(t.shape) == tuple(s['shape']), (what, t.shape, s['shape'])
    f = t.detach().reshape(-1).float().cpu()
    v = f[::s['step']][:s['values'].numel()]
    err = float((v - s['values']).abs().max()) / max(s['absmax'], 1e-30)
    nerr = abs(float(f.double().norm()) - s['norm']) / max(s['norm'], 1e-30)
    assert err <= tol, '%s: sampled rel err %.3e > %.1e' % (what, err, tol)
    assert nerr <= tol, '%s: norm rel err %.3e > %.1e' % (what, nerr, tol)
    return err


def check_sample_l2(t, s, tol, what=''):
    """rel-L2 over the sampled entries (+ norm): the measure for chaotic quantities (tiny-batch gradients)"""
    assert tuple(t.shape) == tuple(s['shape']), (what, t.shape, s['shape'])
    f = t.detach().reshape(-1).float().cpu()
    v = f[::s['step']][:s['values'].numel()].double()
    ref = s['values'].double()
    err = float((v - ref).norm() / ref.norm().clamp_min(1e-30))
    nerr = abs(float(f.double().norm()) - s['norm']) / max(s['norm'], 1e-30)
    assert err <= tol, '%s: sampled rel-L2 err %.3e > %.1e' % (what, err, tol)
    assert nerr <= tol, '%s: norm rel err %.3e > %.1e' % (what, nerr, tol)
    return err
