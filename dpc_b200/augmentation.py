"""On-device clip augmentation behind the reference's transform vocabulary.

The reference builds `transforms.Compose([RandomSizedCrop(...), RandomHorizontalFlip(...), RandomGray(...), ColorJitter(...),
ToTensor(), Normalize()])` (/root/reference/dpc/main.py:115-133) from the classes of /root/reference/utils/augmentation.py and
runs it over PIL images inside 32 DataLoader workers (main.py:307-321).  Here the classes with the same names and constructor
arguments do not touch pixels: each one only draws ITS random decisions -- with the same calls, in the same order, on
Python's `random` and numpy's global generator as the reference class -- into a `ClipPlan`.  `Compose` then folds crop, flips
and resize of a clip into separable resampling tables, packs the per-frame grey / colour-jitter decisions, and ONE CUDA
kernel (csrc/augment.cu, `dpc_augment_clips`) turns the decoded uint8 frames of a whole batch into the float32 block
`[B, num_seq, 3, seq_len, H, W]` that `DPC_RNN.forward` takes -- bit-identical to what the reference's CPU chain produces
from the same frames under the same seeds (tests/test_augment.py against tests/golden/aug_*.pt).

    transform = Compose([RandomSizedCrop(size=128, consistent=True, p=1.0), RandomHorizontalFlip(consistent=True),
                         RandomGray(consistent=False, p=0.5),
                         ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0),
                         ToTensor(), Normalize()])
    block = transform(frames_u8, num_seq=8, seq_len=5)      # frames_u8: CUDA uint8 [B, 40, H, W, 3]

Supported: the two recipes of main.py (k400: RandomSizedCrop incl. its Scale + CenterCrop fallback; ucf101: flip ->
RandomCrop(224) -> Scale((d, d)) NEAREST), RandomGray(consistent=False), ColorJitter(consistent=False), geometric steps
`consistent=True`.  Anything else raises NotImplementedError -- there is no CPU fallback.
"""
import functools
import math
import random

import numpy as np
import torch

from ._lib import lib, ptr

BILINEAR, NEAREST = 'bilinear', 'nearest'
_PREC = 22                                            # Pillow's fixed-point coefficient scale for 8-bit images
_OPS = {'brightness': 0, 'contrast': 1, 'saturation': 2, 'hue': 3}


class ClipPlan:
    """the random decisions of one clip (n_frames frames of W x H pixels)"""

    def __init__(self, n_frames, W, H):
        self.n_frames, self.W, self.H = n_frames, W, H
        self.box = (0, 0, W, H)                       # crop window (x, y, w, h) of the (possibly mirrored) source
        self.mirror_src = False                       # horizontal flip applied before any crop / resize
        self.mirror_out = False                       # horizontal flip applied after the resize
        self.resample = None                          # None | (BILINEAR | NEAREST, (Wo, Ho))
        self.window = None                            # None | (x, y, w, h) kept out of the resized grid
        self.gray = np.full(n_frames, -1, np.int32)
        self.ops = np.full((n_frames, 4), -1, np.int32)
        self.factors = np.zeros((n_frames, 4), np.float32)
        self.hue = np.zeros(n_frames, np.int32)
        self.to_tensor = self.normalize = None

    @property
    def cur_size(self):
        """(w, h) of the image list at this point of the chain"""
        if self.window is not None:
            return self.window[2], self.window[3]
        return self.resample[1] if self.resample is not None else (self.box[2], self.box[3])

    def _geometry_open(self, who):
        if self.resample is not None or self.mirror_out:
            raise NotImplementedError('%s after a resize (or after a flip that followed a crop) is not supported on the '
                                      'device path' % who)


# ---------------------------------------------------------------------------------------------------------------
# the reference's transform classes: same names / arguments, each contributes its random draws to the plan
# ---------------------------------------------------------------------------------------------------------------
class RandomSizedCrop:
    """utils/augmentation.py:147-203"""

    def __init__(self, size, interpolation=BILINEAR, consistent=True, p=1.0):
        if interpolation != BILINEAR or not consistent:
            raise NotImplementedError('RandomSizedCrop: only interpolation=BILINEAR, consistent=True')
        self.size, self.threshold = size, p

    def plan(self, P):
        P._geometry_open('RandomSizedCrop')
        x0, y0, W, H = P.box
        S = self.size
        if not random.random() < self.threshold:
            raise NotImplementedError('RandomSizedCrop with p < 1 (CenterCrop branch)')
        for _ in range(10):
            area = W * H
            target_area = random.uniform(0.5, 1) * area
            aspect_ratio = random.uniform(3. / 4, 4. / 3)
            w = int(round(math.sqrt(target_area * aspect_ratio)))
            h = int(round(math.sqrt(target_area / aspect_ratio)))
            if random.random() < 0.5:
                w, h = h, w
            if w <= W and h <= H:
                x1 = random.randint(0, W - w)
                y1 = random.randint(0, H - h)
                P.box = (x0 + x1, y0 + y1, w, h)
                P.resample = (BILINEAR, (S, S))
                return
        # ten misses (common for 16:9 frames): Scale(size) to the short side, then CenterCrop(size)
        if (W <= H and W == S) or (H <= W and H == S):
            ow, oh = W, H
        elif W < H:
            ow, oh = S, int(S * H / W)
        else:
            ow, oh = int(S * W / H), S
        cx, cy = int(round((ow - S) / 2.)), int(round((oh - S) / 2.))
        if (ow, oh) == (W, H):
            P.box = (x0 + cx, y0 + cy, S, S)
        else:
            P.resample = (BILINEAR, (ow, oh))
            P.window = (cx, cy, S, S)


class RandomCrop:
    """utils/augmentation.py:98-144 (consistent, no flow map)"""

    def __init__(self, size, consistent=True):
        if not consistent:
            raise NotImplementedError('RandomCrop: only consistent=True')
        self.size = (int(size), int(size)) if isinstance(size, (int, float)) else size

    def plan(self, P):
        P._geometry_open('RandomCrop')
        x0, y0, W, H = P.box
        th, tw = self.size
        if W == tw and H == th:
            return
        x1 = random.randint(0, W - tw)
        y1 = random.randint(0, H - th)
        P.box = (x0 + x1, y0 + y1, tw, th)


class Scale:
    """utils/augmentation.py:20-41 with a (w, h) size"""

    def __init__(self, size, interpolation=NEAREST):
        if isinstance(size, int) or interpolation not in (NEAREST, BILINEAR):
            raise NotImplementedError('Scale: only a (w, h) size with NEAREST or BILINEAR')
        self.size, self.interpolation = tuple(size), interpolation

    def plan(self, P):
        P._geometry_open('Scale')
        P.resample = (self.interpolation, self.size)


class RandomHorizontalFlip:
    """utils/augmentation.py:206-232"""

    def __init__(self, consistent=True, command=None):
        if not consistent:
            raise NotImplementedError('RandomHorizontalFlip: only consistent=True')
        self.threshold = 0 if command == 'left' else (1 if command == 'right' else 0.5)

    def plan(self, P):
        if random.random() < self.threshold:
            if P.resample is None and P.box == (0, 0, P.W, P.H):
                P.mirror_src = not P.mirror_src       # nothing geometric yet: the frame itself is mirrored
            else:
                P.mirror_out = not P.mirror_out       # mirrors the grid produced so far


class RandomGray:
    """utils/augmentation.py:235-261: 'a channel splitting, not strictly grayscale'"""

    def __init__(self, consistent=True, p=0.5):
        if consistent:
            raise NotImplementedError('RandomGray: only consistent=False (as in main.py)')
        self.p = p

    def plan(self, P):
        for f in range(P.n_frames):
            if random.random() < self.p:
                P.gray[f] = int(np.random.choice(3))


class ColorJitter:
    """utils/augmentation.py:264-355"""

    def __init__(self, brightness=0, contrast=0, saturation=0, hue=0, consistent=False, p=1.0):
        if consistent:
            raise NotImplementedError('ColorJitter: only consistent=False (as in main.py)')
        self.brightness = self._check_input(brightness, 'brightness')
        self.contrast = self._check_input(contrast, 'contrast')
        self.saturation = self._check_input(saturation, 'saturation')
        self.hue = self._check_input(hue, 'hue', center=0, bound=(-0.5, 0.5), clip_first_on_zero=False)
        self.threshold = p

    @staticmethod
    def _check_input(value, name, center=1, bound=(0, float('inf')), clip_first_on_zero=True):
        if isinstance(value, (int, float)):
            if value < 0:
                raise ValueError('If {} is a single number, it must be non negative.'.format(name))
            value = [center - value, center + value]
            if clip_first_on_zero:
                value[0] = max(value[0], 0)
        elif isinstance(value, (tuple, list)) and len(value) == 2:
            if not bound[0] <= value[0] <= value[1] <= bound[1]:
                raise ValueError('{} values should be between {}'.format(name, bound))
        else:
            raise TypeError('{} should be a single number or a list/tuple with lenght 2.'.format(name))
        return None if value[0] == value[1] == center else value

    def plan(self, P):
        if not random.random() < self.threshold:
            return
        for f in range(P.n_frames):
            chain = []                                # get_params: the four uniforms first, then the shuffle
            for name, rng in (('brightness', self.brightness), ('contrast', self.contrast),
                              ('saturation', self.saturation), ('hue', self.hue)):
                if rng is not None:
                    chain.append((_OPS[name], random.uniform(rng[0], rng[1])))
            random.shuffle(chain)
            for i, (op, factor) in enumerate(chain):
                P.ops[f, i] = op
                P.factors[f, i] = factor              # Pillow's blend takes a C float
                if op == _OPS['hue']:
                    if not -0.5 <= factor <= 0.5:
                        raise ValueError('hue_factor ({}) is not in [-0.5, 0.5].'.format(factor))
                    P.hue[f] = int(np.int32(factor * 255).astype(np.uint8))      # torchvision adjust_hue


class ToTensor:
    """utils/augmentation.py:373-376"""

    def plan(self, P):
        P.to_tensor = True


class Normalize:
    """utils/augmentation.py:378-384"""

    def __init__(self, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]):
        self.mean, self.std = list(mean), list(std)

    def plan(self, P):
        if not P.to_tensor:
            raise NotImplementedError('Normalize before ToTensor')
        P.normalize = (self.mean, self.std)


# ---------------------------------------------------------------------------------------------------------------
# resampling tables (Pillow's Resample.c coefficients, in double precision exactly as the C code computes them)
# ---------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=4096)
def _bilinear_table(n_in, n_out):
    scale = n_in / n_out
    fscale = scale if scale > 1.0 else 1.0
    support = 1.0 * fscale
    K = int(math.ceil(support)) * 2 + 1
    center = (np.arange(n_out, dtype=np.float64) + 0.5) * scale
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)            # C (int) casts truncate; operands are >= -0.5
    hi = np.minimum((center + support + 0.5).astype(np.int64), n_in)
    cnt = hi - lo
    j = np.arange(K, dtype=np.float64)[None, :]
    arg = np.abs((j + lo[:, None] - center[:, None] + 0.5) * (1.0 / fscale))
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(np.arange(K)[None, :] < cnt[:, None], w, 0.0)
    ww = np.zeros(n_out, np.float64)
    for k in range(K):                                # left-to-right accumulation, as the C loop
        ww = ww + w[:, k]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    coef = (0.5 + w * (1 << _PREC)).astype(np.int64)  # weights are non-negative: (int)(0.5 + k * 2^22)
    coef = np.where(np.arange(K)[None, :] < cnt[:, None], coef, 0)
    return lo.astype(np.int32), cnt.astype(np.int32), coef.astype(np.int32)


@functools.lru_cache(maxsize=256)
def _nearest_table(n_in, n_out):
    scale = n_in / n_out
    xo = scale * 0.5
    idx = np.zeros(n_out, np.int32)
    for x in range(n_out):                            # Pillow accumulates the source coordinate
        idx[x] = min(max(int(math.floor(xo)), 0), n_in - 1)
        xo += scale
    return idx, np.ones(n_out, np.int32), np.full((n_out, 1), 1 << _PREC, np.int32)


@functools.lru_cache(maxsize=256)
def _identity_table(n):
    return np.arange(n, dtype=np.int32), np.ones(n, np.int32), np.full((n, 1), 1 << _PREC, np.int32)


def clip_tables(P):
    """(xstart, xcount, xcoef, xstep), (ystart, ycount, ycoef) of a plan, in source-frame coordinates"""
    x0, y0, w, h = P.box
    if P.resample is None:
        tx, ty = _identity_table(w), _identity_table(h)
    else:
        kind, (Wo, Ho) = P.resample
        make = _bilinear_table if kind == BILINEAR else _nearest_table
        tx, ty = make(w, Wo), make(h, Ho)
    if P.window is not None:
        wx, wy, ww, wh = P.window
        tx = tuple(a[wx:wx + ww] for a in tx)
        ty = tuple(a[wy:wy + wh] for a in ty)
    xs, xc, xk = tx
    ys, yc, yk = ty
    if P.mirror_out:
        xs, xc, xk = xs[::-1], xc[::-1], xk[::-1]
    if P.mirror_src:                                  # column u of the mirrored source is column W - 1 - u of the frame
        return (P.W - 1 - (x0 + xs), xc, xk, -1), (y0 + ys, yc, yk)
    return (x0 + xs, xc, xk, 1), (y0 + ys, yc, yk)


class Compose:
    """the transform chain; callable on a batch of decoded clips"""

    def __init__(self, transforms):
        self.transforms = list(transforms)
        for t in self.transforms:
            if not hasattr(t, 'plan'):
                raise NotImplementedError('%s has no device implementation' % type(t).__name__)

    def plan(self, n_frames, W, H):
        """draw one clip's random decisions (consumes `random` / `numpy.random` exactly as the reference chain would)"""
        P = ClipPlan(n_frames, W, H)
        for t in self.transforms:
            t.plan(P)
        if not P.to_tensor or P.normalize is None:
            raise NotImplementedError('the device chain ends with ToTensor() and Normalize()')
        return P

    @staticmethod
    def pack(plans):
        """plans of one batch -> (tables int32 [B, L], frame_params int32 [B, F, 10], (Wo, Ho), K)"""
        tabs = [clip_tables(P) for P in plans]
        Wo, Ho = plans[0].cur_size
        K = max(max(tx[2].shape[1], ty[2].shape[1]) for tx, ty in tabs)
        L = (Wo + Ho) * (2 + K) + 1
        tables = np.zeros((len(plans), L), np.int32)
        fpar = np.zeros((len(plans), plans[0].n_frames, 10), np.int32)
        for b, (P, (tx, ty)) in enumerate(zip(plans, tabs)):
            if P.cur_size != (Wo, Ho) or P.n_frames != plans[0].n_frames:
                raise ValueError('clips of one batch must share the output size and frame count')
            xs, xc, xk, xstep = tx
            ys, yc, yk = ty
            o = 0
            tables[b, o:o + Wo] = xs; o += Wo
            tables[b, o:o + Wo] = xc; o += Wo
            tables[b, o:o + Wo * K].reshape(Wo, K)[:, :xk.shape[1]] = xk; o += Wo * K
            tables[b, o:o + Ho] = ys; o += Ho
            tables[b, o:o + Ho] = yc; o += Ho
            tables[b, o:o + Ho * K].reshape(Ho, K)[:, :yk.shape[1]] = yk; o += Ho * K
            tables[b, o] = xstep
            fpar[b, :, 0] = P.gray
            fpar[b, :, 1:5] = P.ops
            fpar[b, :, 5:9] = P.factors.view(np.int32)
            fpar[b, :, 9] = P.hue
        return tables, fpar, (Wo, Ho), K

    def prepare(self, plans, device):
        """pack the plans of one batch and start their upload (pinned staging, asynchronous on the current stream); the
        result can be handed to __call__ later -- e.g. prepared under the previous training step"""
        tables, fpar, (Wo, Ho), K = self.pack(plans)
        with torch.cuda.device(device):
            t_d = torch.from_numpy(tables).pin_memory().to(device, non_blocking=True)
            f_d = torch.from_numpy(fpar).pin_memory().to(device, non_blocking=True)
        mean = np.asarray(plans[0].normalize[0], np.float32)
        std = np.asarray(plans[0].normalize[1], np.float32)
        return dict(tables=t_d, frame_params=f_d, size=(Wo, Ho), K=K, mean=mean, std=std, B=len(plans), F=plans[0].n_frames,
                    src=(plans[0].W, plans[0].H))

    def __call__(self, frames, num_seq, seq_len, plans=None, out=None, prepared=None):
        """frames: CUDA uint8 [B, F, H, W, 3] (or [F, H, W, 3]) decoded RGB frames, F = num_seq * seq_len
        -> float32 block [B, num_seq, 3, seq_len, Ho, Wo] (dataset_3d.py:108-112 layout, batched)"""
        if frames.dtype != torch.uint8 or frames.dim() not in (4, 5) or frames.shape[-1] != 3:
            raise ValueError('expected uint8 frames [B, F, H, W, 3], got %s %s' % (frames.dtype, tuple(frames.shape)))
        if not frames.is_cuda:
            raise RuntimeError('dpc_b200 has no CPU path: frames must be a CUDA tensor')
        if frames.dim() == 4:
            frames = frames[None]
        frames = frames.contiguous()
        B, F, H, W, _ = frames.shape
        if F != num_seq * seq_len:
            raise ValueError('num_seq * seq_len = %d but the clips have %d frames' % (num_seq * seq_len, F))
        if prepared is None:
            if plans is None:
                plans = [self.plan(F, W, H) for _ in range(B)]
            prepared = self.prepare(plans, frames.device)
        if (prepared['B'], prepared['F'], prepared['src']) != (B, F, (W, H)):
            raise ValueError('the prepared batch is for %s clips x %s frames of %s, the frames are %s'
                             % (prepared['B'], prepared['F'], prepared['src'], tuple(frames.shape)))
        Wo, Ho = prepared['size']
        with torch.cuda.device(frames.device):
            if out is None:
                out = torch.empty(B, num_seq, 3, seq_len, Ho, Wo, device=frames.device)
            elif tuple(out.shape) != (B, num_seq, 3, seq_len, Ho, Wo) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError('out must be a contiguous float32 [%d, %d, 3, %d, %d, %d]' % (B, num_seq, seq_len, Ho, Wo))
            lib().augment_clips(ptr(frames), ptr(prepared['tables']), ptr(prepared['frame_params']),
                                prepared['mean'].ctypes.data, prepared['std'].ctypes.data, ptr(out), B, F, H, W, Ho, Wo,
                                prepared['K'], num_seq, seq_len, torch.cuda.current_stream().cuda_stream)
        return out


def k400_transform(img_dim):
    """main.py:125-133"""
    return Compose([RandomSizedCrop(size=img_dim, consistent=True, p=1.0), RandomHorizontalFlip(consistent=True),
                    RandomGray(consistent=False, p=0.5),
                    ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0), ToTensor(), Normalize()])


def ucf101_transform(img_dim):
    """main.py:115-124"""
    return Compose([RandomHorizontalFlip(consistent=True), RandomCrop(size=224, consistent=True),
                    Scale(size=(img_dim, img_dim)), RandomGray(consistent=False, p=0.5),
                    ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0), ToTensor(), Normalize()])
