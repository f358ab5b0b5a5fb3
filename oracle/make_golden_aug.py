"""Generate tests/golden/aug_*.pt by running the UNMODIFIED reference augmentation classes
(/root/reference/utils/augmentation.py, composed exactly as /root/reference/dpc/main.py:115-133 and reshaped as
/root/reference/dpc/dataset_3d.py:108-112) on synthetic decoded frames, under Pillow + torchvision of this image.

Test infrastructure; runs only in the build container.  Usage:  python oracle/make_golden_aug.py
Each fixture stores the case (recipe, frame geometry, seeds) and a fingerprint of the reference's output block
(shape, sha256 of the float32 bytes, strided sample) -- the frames are regenerated from the seed by
oracle.aug_oracle.make_frames, the random parameters by seeding `random` / `numpy.random` identically.
"""
import collections
import collections.abc
import os
import random
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/utils')

# environment shim, not a change to the reference: Scale.__init__ (augmentation.py:22) names collections.Iterable, which
# Python >= 3.10 only has under collections.abc
if not hasattr(collections, 'Iterable'):
    collections.Iterable = collections.abc.Iterable

import augmentation as ref_aug          # noqa: E402  (the reference itself)
from oracle import aug_oracle as A      # noqa: E402

CASES = [
    # name, recipe, W, H, num_seq, seq_len, img_dim, frame seed, rng seed
    ('aug_k400_small_s1', 'k400', 100, 75, 2, 3, 64, 1, 101),
    ('aug_k400_small_s2', 'k400', 100, 75, 2, 3, 64, 2, 102),
    ('aug_k400_full_s3', 'k400', 200, 150, 8, 5, 128, 3, 103),
    ('aug_k400_wide_s4', 'k400', 267, 150, 8, 5, 128, 4, 104),
    ('aug_k400_224_s5', 'k400', 340, 256, 8, 5, 224, 5, 105),
    ('aug_ucf101_s6', 'ucf101', 341, 256, 8, 5, 128, 6, 106),
    ('aug_ucf101_s7', 'ucf101', 320, 240, 4, 5, 128, 7, 107),
]


def reference_transform(recipe, img_dim):
    if recipe == 'ucf101':                                # main.py:115-124
        return transforms.Compose([
            ref_aug.RandomHorizontalFlip(consistent=True),
            ref_aug.RandomCrop(size=224, consistent=True),
            ref_aug.Scale(size=(img_dim, img_dim)),
            ref_aug.RandomGray(consistent=False, p=0.5),
            ref_aug.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0),
            ref_aug.ToTensor(),
            ref_aug.Normalize()])
    return transforms.Compose([                           # main.py:125-133
        ref_aug.RandomSizedCrop(size=img_dim, consistent=True, p=1.0),
        ref_aug.RandomHorizontalFlip(consistent=True),
        ref_aug.RandomGray(consistent=False, p=0.5),
        ref_aug.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0),
        ref_aug.ToTensor(),
        ref_aug.Normalize()])


def reference_block(recipe, frames, num_seq, seq_len, img_dim, seed):
    random.seed(seed)
    np.random.seed(seed)
    seq = [Image.fromarray(f, 'RGB') for f in frames]
    t_seq = reference_transform(recipe, img_dim)(seq)
    (C, H, W) = t_seq[0].size()
    t_seq = torch.stack(t_seq, 0)                         # dataset_3d.py:108-112
    return t_seq.view(num_seq, seq_len, C, H, W).transpose(1, 2).contiguous().numpy()


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, recipe, W, H, N, SL, S, fseed, rseed in CASES:
        frames = A.make_frames(fseed, N * SL, H, W)
        block = reference_block(recipe, frames, N, SL, S, rseed)
        # the oracle must reproduce the reference bit for bit before the fixture is written
        random.seed(rseed)
        np.random.seed(rseed)
        plan = (A.plan_ucf101 if recipe == 'ucf101' else A.plan_k400)(N * SL, W, H, S)
        mine, _ = A.augment_clip(frames, plan, N, SL)
        same = np.array_equal(mine.view(np.uint32), block.view(np.uint32))
        fp = A.fingerprint(block)
        fp.update(name=name, recipe=recipe, W=W, H=H, num_seq=N, seq_len=SL, img_dim=S, frame_seed=fseed, rng_seed=rseed,
                  pillow=Image.__version__ if hasattr(Image, '__version__') else '', box=plan.box,
                  flip=(plan.flip_src, plan.flip_out))
        fp['sample'] = torch.from_numpy(fp['sample'])
        torch.save(fp, os.path.join(out_dir, name + '.pt'))
        print('%-20s block %s  oracle == reference: %s   box %s flip %s' % (name, block.shape, same, plan.box,
                                                                          (plan.flip_src, plan.flip_out)))
        assert same, name


if __name__ == '__main__':
    main()
