"""GPU: the on-device augmentation kernel (csrc/augment.cu through dpc_b200.augmentation.Compose) against the oracle and the
reference-generated fingerprints -- bit-exact -- plus batch-level properties at the benchmarked size."""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import aug_oracle as A

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'aug_*.pt')))


def _transform(recipe, img_dim):
    from dpc_b200 import augmentation as D
    return D.ucf101_transform(img_dim) if recipe == 'ucf101' else D.k400_transform(img_dim)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_kernel_reproduces_reference_chain(path):
    fx = torch.load(path, weights_only=False)
    N, SL = fx['num_seq'], fx['seq_len']
    frames = A.make_frames(fx['frame_seed'], N * SL, fx['H'], fx['W'])
    random.seed(fx['rng_seed'])
    np.random.seed(fx['rng_seed'])
    block = _transform(fx['recipe'], fx['img_dim'])(torch.from_numpy(frames).cuda(), N, SL)
    assert block.shape == (1, N, 3, SL, fx['img_dim'], fx['img_dim'])
    got = A.fingerprint(block[0].cpu().numpy())
    if got['sha256'] != fx['sha256']:                 # diagnose against the oracle before failing
        random.seed(fx['rng_seed'])
        np.random.seed(fx['rng_seed'])
        plan = (A.plan_ucf101 if fx['recipe'] == 'ucf101' else A.plan_k400)(N * SL, fx['W'], fx['H'], fx['img_dim'])
        ref, _ = A.augment_clip(frames, plan, N, SL)
        d = np.argwhere(ref != block[0].cpu().numpy())
        pytest.fail('%d of %d values differ from the oracle, first at %s (frame ops %s)'
                    % (len(d), ref.size, d[:3].tolist(), plan.jitter[int(d[0][0]) * SL + int(d[0][2])]))
    assert np.array_equal(got['sample'], fx['sample'].numpy())


def test_batch_matches_oracle_per_clip_and_is_deterministic():
    """a batch of clips with different crops / flips / orders: every clip equals the oracle's single-clip result (bit-exact),
    the batched launch equals clip-by-clip launches, and re-running the same plans reproduces the block"""
    from dpc_b200 import augmentation as D
    B, N, SL, W, H, S = 6, 4, 5, 200, 150, 128
    frames = np.stack([A.make_frames(50 + b, N * SL, H, W) for b in range(B)])
    tr = D.k400_transform(S)
    random.seed(9)
    np.random.seed(9)
    plans = [tr.plan(N * SL, W, H) for _ in range(B)]
    fd = torch.from_numpy(frames).cuda()
    block = tr(fd, N, SL, plans=plans)
    again = tr(fd, N, SL, plans=plans)
    assert torch.equal(block, again)
    random.seed(9)
    np.random.seed(9)
    for b in range(B):
        ref, _ = A.augment_clip(frames[b], A.plan_k400(N * SL, W, H, S), N, SL)
        assert np.array_equal(block[b].cpu().numpy().view(np.uint32), ref.view(np.uint32)), b
        single = tr(fd[b], N, SL, plans=[plans[b]])
        assert torch.equal(single[0], block[b])


def test_full_batch_properties():
    """BASELINE config-2 sized batch (128 clips x 40 frames -> [128, 8, 3, 5, 128, 128]): finite, inside the normalised
    uint8 range, grey frames have identical channels up to the per-channel normalisation, and a sample of clips is
    bit-identical to the same clips run alone"""
    from dpc_b200 import augmentation as D
    B, N, SL, W, H, S = 128, 8, 5, 200, 150, 128
    g = torch.Generator(device='cuda').manual_seed(3)
    frames = torch.randint(0, 256, (B, N * SL, H, W, 3), dtype=torch.uint8, device='cuda', generator=g)
    tr = D.k400_transform(S)
    random.seed(10)
    np.random.seed(10)
    plans = [tr.plan(N * SL, W, H) for _ in range(B)]
    block = tr(frames, N, SL, plans=plans)
    torch.cuda.synchronize()
    assert block.shape == (B, N, 3, SL, S, S) and torch.isfinite(block).all()
    mean = torch.tensor(A.MEAN, device='cuda').view(1, 1, 3, 1, 1, 1)
    std = torch.tensor(A.STD, device='cuda').view(1, 1, 3, 1, 1, 1)
    u8 = (block * std + mean) * 255
    assert float((u8 - u8.round()).abs().max()) < 1e-3 and float(u8.min()) > -1e-3 and float(u8.max()) < 255 + 1e-3
    for b in (0, 77, 127):
        assert torch.equal(tr(frames[b], N, SL, plans=[plans[b]])[0], block[b])
    b, f = next((b, f) for b in range(B) for f in range(N * SL) if plans[b].gray[f] >= 0 and plans[b].ops[f, 0] < 0) \
        if any(p.ops[:, 0].min() < 0 for p in plans) else (None, None)
    if b is not None:
        px = u8[b, f // SL, :, f % SL].round()
        assert torch.equal(px[0], px[1]) and torch.equal(px[1], px[2])


def _random_case(rng, recipe):
    if recipe == 'ucf101':
        W, H = rng.randint(224, 300), rng.randint(224, 280)
        S = rng.choice([64, 96, 112, 128])
    else:
        S = rng.choice([32, 48, 64, 96, 128])
        W, H = rng.randint(S // 2 + 8, 3 * S), rng.randint(S // 2 + 8, 3 * S)      # up- and down-scaling, any aspect
    return W, H, S


@pytest.mark.parametrize('recipe', ['k400', 'ucf101'])
def test_random_geometries_match_oracle(recipe):
    """randomised frame sizes / output sizes (up-scaling, strong down-scaling with wide tap windows, the Scale + CenterCrop
    fallback for extreme aspect ratios, flips on both sides of the resize): every case bit-exact against the oracle"""
    from dpc_b200 import augmentation as D
    rng = random.Random(2024 if recipe == 'k400' else 2025)
    for case in range(10 if recipe == 'k400' else 4):
        W, H, S = _random_case(rng, recipe)
        N, SL = 2, 2
        frames = A.make_frames(900 + case, N * SL, H, W)
        seed = 5000 + case
        random.seed(seed)
        np.random.seed(seed)
        tr = D.ucf101_transform(S) if recipe == 'ucf101' else D.k400_transform(S)
        got = tr(torch.from_numpy(frames).cuda(), N, SL)[0].cpu().numpy()
        random.seed(seed)
        np.random.seed(seed)
        plan = (A.plan_ucf101 if recipe == 'ucf101' else A.plan_k400)(N * SL, W, H, S)
        ref, _ = A.augment_clip(frames, plan, N, SL)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (recipe, case, W, H, S, plan.box, plan.resize)


def test_odd_output_size_and_bilinear_scale():
    """an output whose pixel count is not a multiple of 4 (scalar store path) and Scale(..., BILINEAR)"""
    from dpc_b200 import augmentation as D
    W, H, N, SL = 90, 70, 1, 3
    frames = A.make_frames(77, N * SL, H, W)
    tr = D.Compose([D.RandomCrop(size=(50, 61)), D.Scale(size=(37, 29), interpolation=D.BILINEAR), D.RandomHorizontalFlip(),
                    D.RandomGray(consistent=False, p=0.5), D.ColorJitter(0.4, 0.3, 0.2, 0.1), D.ToTensor(), D.Normalize()])
    random.seed(31)
    np.random.seed(31)
    got = tr(torch.from_numpy(frames).cuda(), N, SL)[0].cpu().numpy()
    random.seed(31)
    np.random.seed(31)
    plan = A.ClipPlan(N * SL, W, H)
    A.plan_random_crop(plan, (50, 61))
    plan.resize = ('bilinear', (37, 29))
    A.plan_flip(plan)
    A.plan_gray(plan)
    A.plan_jitter(plan, 0.4, 0.3, 0.2, 0.1)
    ref, _ = A.augment_clip(frames, plan, N, SL)
    assert got.shape == (1, 3, 3, 29, 37)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize('hue', [-0.25, 0.1037, 0.5])
def test_hue_rotation_all_colours(hue):
    """every one of the 2^24 RGB colours through the kernel's HSV round trip (1024 frames of 128 x 128, identity geometry,
    hue as the only jitter operation) against the oracle -- which tests/test_aug_oracle.py pins against Pillow on the same
    exhaustive cube"""
    from dpc_b200 import augmentation as D
    r = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(r, r, r, indexing='ij'), -1).reshape(1024, 128, 128, 3)
    tr = D.Compose([D.ColorJitter(hue=(hue, hue), p=1.0), D.ToTensor(), D.Normalize()])
    got = tr(torch.from_numpy(cube).cuda(), 1024, 1)
    torch.cuda.synchronize()
    ref = A.adjust(cube.reshape(4096, 4096, 3), A.OP_HUE, hue).reshape(1024, 128, 128, 3)
    lut = A.normalize(np.arange(256, dtype=np.uint8).reshape(256, 1, 1).repeat(3, -1))[..., 0]     # [3, 256] bytes -> floats
    want = np.stack([lut[c][ref[..., c]] for c in range(3)], 1)                                   # [1024, 3, 128, 128]
    g = got[0, :, :, 0].cpu().numpy()                                                             # [1024, 3, 128, 128]
    bad = np.argwhere(g.view(np.uint32) != want.view(np.uint32))
    assert len(bad) == 0, (len(bad), bad[:3].tolist())
