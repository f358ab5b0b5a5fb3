"""CPU: the augmentation oracle (oracle/aug_oracle.py) against the live Pillow / torchvision of this image and against the
fingerprints of the UNMODIFIED reference chain (tests/golden/aug_*.pt, oracle/make_golden_aug.py); the product's host side
(dpc_b200/augmentation.py: random draws + resampling tables) against the oracle's under the same seeds."""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import aug_oracle as A

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'aug_*.pt')))


def _plan(fx):
    random.seed(fx['rng_seed'])
    np.random.seed(fx['rng_seed'])
    make = A.plan_ucf101 if fx['recipe'] == 'ucf101' else A.plan_k400
    return make(fx['num_seq'] * fx['seq_len'], fx['W'], fx['H'], fx['img_dim'])


def test_fixtures_exist():
    assert len(GOLDEN) >= 7


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_reproduces_reference_chain(path):
    """bit-exact: sha256 of the float32 block equals the one the unmodified reference classes produced"""
    fx = torch.load(path, weights_only=False)
    frames = A.make_frames(fx['frame_seed'], fx['num_seq'] * fx['seq_len'], fx['H'], fx['W'])
    plan = _plan(fx)
    assert tuple(plan.box) == tuple(fx['box']) and (plan.flip_src, plan.flip_out) == tuple(fx['flip'])
    block, _ = A.augment_clip(frames, plan, fx['num_seq'], fx['seq_len'])
    fp = A.fingerprint(block)
    assert fp['shape'] == tuple(fx['shape'])
    assert np.array_equal(fp['sample'], fx['sample'].numpy())
    assert fp['sha256'] == fx['sha256']


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_product_plan_and_tables_match_oracle(path):
    """dpc_b200.augmentation draws the same decisions and builds the same fixed-point tables as the oracle"""
    from dpc_b200 import augmentation as D
    fx = torch.load(path, weights_only=False)
    F = fx['num_seq'] * fx['seq_len']
    plan = _plan(fx)
    random.seed(fx['rng_seed'])
    np.random.seed(fx['rng_seed'])
    tr = D.ucf101_transform(fx['img_dim']) if fx['recipe'] == 'ucf101' else D.k400_transform(fx['img_dim'])
    P = tr.plan(F, fx['W'], fx['H'])
    (xs, xc, xk, xstep), (ys, yc, yk, _) = plan.tables()
    (pxs, pxc, pxk, pxstep), (pys, pyc, pyk) = D.clip_tables(P)
    assert pxstep == xstep
    for a, b in ((pxs, xs), (pxc, xc), (pys, ys), (pyc, yc)):
        assert np.array_equal(np.asarray(a, np.int64), np.asarray(b, np.int64))
    for a, b, c in ((pxk, xk, xc), (pyk, yk, yc)):
        K = min(a.shape[1], b.shape[1])
        assert np.array_equal(a[:, :K].astype(np.int64), b[:, :K]) and not a[:, K:].any() and not b[:, K:].any()
        assert not (np.arange(a.shape[1])[None, :] >= np.asarray(c)[:, None])[a != 0].any()      # no weight beyond the count
    assert list(P.gray) == list(plan.gray)
    for f in range(F):
        ops = [(int(o), float(np.float32(x))) for o, x in zip(P.ops[f], P.factors[f]) if o >= 0]
        assert ops == [(o, float(np.float32(x))) for o, x in plan.jitter[f]]
        hue = [x for o, x in plan.jitter[f] if o == A.OP_HUE]
        assert int(P.hue[f]) == (A.hue_shift_byte(hue[0]) if hue else 0)
    tables, fpar, (Wo, Ho), K = tr.pack([P])
    assert (Wo, Ho) == plan.out_size() and tables.shape == (1, (Wo + Ho) * (2 + K) + 1) and fpar.shape == (1, F, 10)


def test_unsupported_chains_fail_loudly():
    from dpc_b200 import augmentation as D
    with pytest.raises(NotImplementedError):
        D.RandomGray(consistent=True)
    with pytest.raises(NotImplementedError):
        D.ColorJitter(brightness=0.5, consistent=True)
    with pytest.raises(NotImplementedError):
        D.Compose([D.RandomSizedCrop(64), D.ToTensor()]).plan(4, 100, 80)
    with pytest.raises(RuntimeError, match='no CPU path'):
        D.k400_transform(64)(torch.zeros(1, 4, 80, 100, 3, dtype=torch.uint8), 2, 2)
    with pytest.raises(ValueError):
        D.ColorJitter(hue=(-0.7, 0.7))


# ---- the third-party arithmetic the reference delegates to, against the live libraries of this image -----------------
PIL = pytest.importorskip('PIL')


def _pil(img):
    from PIL import Image
    return Image.fromarray(img, 'RGB')


@pytest.mark.parametrize('W,H,S', [(200, 150, 128), (171, 128, 128), (90, 77, 64), (64, 64, 64), (300, 130, 128), (50, 40, 64),
                                   (267, 150, 128), (150, 150, 224)])
def test_bilinear_resize_matches_pillow(W, H, S):
    from PIL import Image
    img = np.random.default_rng(W * 1000 + H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.array(_pil(img).resize((S, S), Image.BILINEAR))
    xs, xc, xk = A.resample_coeffs(W, S)
    ys, yc, yk = A.resample_coeffs(H, S)
    assert np.array_equal(A.apply_tables(img, (xs, xc, xk, 1), (ys, yc, yk, 1)), ref)


@pytest.mark.parametrize('W,H,Wo,Ho', [(224, 224, 128, 128), (224, 224, 96, 96), (100, 80, 64, 64), (224, 224, 100, 100)])
def test_nearest_resize_matches_pillow(W, H, Wo, Ho):
    from PIL import Image
    img = np.random.default_rng(W + H + Wo).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.array(_pil(img).resize((Wo, Ho), Image.NEAREST))
    xs, xc, xk = A.nearest_coeffs(W, Wo)
    ys, yc, yk = A.nearest_coeffs(H, Ho)
    assert np.array_equal(A.apply_tables(img, (xs, xc, xk, 1), (ys, yc, yk, 1)), ref)


def test_blend_matches_pillow_on_all_byte_pairs():
    from PIL import Image
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256, 1).repeat(3, -1)
    b = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(256, 256, 1).repeat(3, -1)
    rng = random.Random(5)
    for f in [0.0, 0.5, 1.0, 1.5, 0.50001, 1.4999] + [rng.uniform(0.5, 1.5) for _ in range(24)]:
        assert np.array_equal(A.blend(a, b, f), np.array(Image.blend(_pil(a), _pil(b), f))), f


def test_hsv_round_trip_matches_pillow_on_all_colours():
    from PIL import Image
    r = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(r, r, r, indexing='ij'), -1).reshape(4096, 4096, 3)
    assert np.array_equal(A.rgb_to_hsv(cube), np.array(_pil(cube).convert('HSV')))
    assert np.array_equal(A.hsv_to_rgb(cube), np.array(Image.fromarray(cube, 'HSV').convert('RGB')))


def test_jitter_steps_and_normalise_match_torchvision():
    F = pytest.importorskip('torchvision.transforms.functional')
    img = np.random.default_rng(1).integers(0, 256, (96, 80, 3), dtype=np.uint8)
    img[:10, :10] = 255
    img[20:30, :10] = 0
    pim = _pil(img)
    rng = random.Random(6)
    for f in [0.5, 1.0, 1.5] + [rng.uniform(0.5, 1.5) for _ in range(8)]:
        assert np.array_equal(np.array(F.adjust_brightness(pim, f)), A.adjust(img, A.OP_BRIGHTNESS, f))
        assert np.array_equal(np.array(F.adjust_contrast(pim, f)), A.adjust(img, A.OP_CONTRAST, f))
        assert np.array_equal(np.array(F.adjust_saturation(pim, f)), A.adjust(img, A.OP_SATURATION, f))
    for f in [-0.25, 0.0, 0.25] + [rng.uniform(-0.25, 0.25) for _ in range(8)]:
        assert np.array_equal(np.array(F.adjust_hue(pim, f)), A.adjust(img, A.OP_HUE, f))
    x = F.normalize(F.to_tensor(pim), A.MEAN, A.STD).numpy()
    assert np.array_equal(x.view(np.uint32), A.normalize(img).view(np.uint32))


def test_vectorised_tables_match_scalar_restatement_over_random_sizes():
    """dpc_b200.augmentation._bilinear_table / _nearest_table (numpy-vectorised, what the product uploads) against the oracle's
    scalar restatement of Pillow's coefficient code over 400 random (input, output) sizes incl. extreme ratios"""
    from dpc_b200 import augmentation as D
    rng = random.Random(77)
    sizes = [(rng.randint(1, 700), rng.randint(1, 260)) for _ in range(380)] + [(1, 1), (1, 64), (700, 1), (256, 256), (2, 255),
                                                                                 (1000, 8), (17, 224), (224, 17)] + \
            [(n, n + d) for n in (31, 64, 127) for d in (-1, 1)] + [(rng.randint(200, 400), 128) for _ in range(6)]
    for n_in, n_out in sizes:
        xs, xc, xk = A.resample_coeffs(n_in, n_out)
        ps, pc, pk = D._bilinear_table(n_in, n_out)
        assert np.array_equal(ps, xs) and np.array_equal(pc, xc), (n_in, n_out)
        assert pk.shape == xk.shape and np.array_equal(pk.astype(np.int64), xk), (n_in, n_out)
        assert int(pk.sum(1).min()) > 0 and abs(int(pk.sum(1).max()) - (1 << 22)) <= pk.shape[1]      # weights sum to ~1.0
        ns, nc, nk = A.nearest_coeffs(n_in, n_out)
        qs, qc, qk = D._nearest_table(n_in, n_out)
        assert np.array_equal(qs, ns) and np.array_equal(qc, nc) and np.array_equal(qk.astype(np.int64), nk), (n_in, n_out)
