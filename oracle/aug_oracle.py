"""CPU restatement (numpy, integer / IEEE-float exact) of the reference's clip augmentation pipeline.

TEST INFRASTRUCTURE ONLY: imported by tests/ and by oracle/make_golden.py, never by the product path
(dpc_b200/augmentation.py samples its own parameters and runs the CUDA kernel).

What is restated, and from where:
  * the transform classes and THEIR RANDOM-NUMBER CONSUMPTION ORDER: /root/reference/utils/augmentation.py
      RandomSizedCrop :147-203, RandomCrop :98-144, Scale :20-41, CenterCrop :44-58, RandomHorizontalFlip :206-232,
      RandomGray :235-261, ColorJitter :264-355, ToTensor :373-376, Normalize :378-384;
    the two recipes of /root/reference/dpc/main.py:115-133 (ucf101, k400) and the frame -> block reshuffle of
    /root/reference/dpc/dataset_3d.py:108-112;
  * the pixel arithmetic those classes delegate to THIRD-PARTY code that is not in /root/reference:
      Pillow (this image: 12.2.0)  Image.resize BILINEAR / NEAREST (libImaging Resample.c: separable convolution,
        support scaled by the down-scale factor, 22-bit fixed-point coefficients, horizontal pass then vertical pass
        with a uint8 intermediate), Image.blend (Blend.c: float32 interpolation, truncation), convert("L") (Convert.c:
        (19595 R + 38470 G + 7471 B + 0x8000) >> 16), convert("HSV") / back (Convert.c, after colorsys.py),
        ImageEnhance.Brightness / Contrast / Color, ImageStat mean;
      torchvision (0.26) transforms.functional adjust_brightness / contrast / saturation / hue (PIL branch), to_tensor
        (uint8 -> float32 .div(255)) and normalize ((x - mean) / std in float32).
    Their published algorithms are restated below; parity is pinned by tests/test_aug_oracle.py against the live Pillow /
    torchvision of this image, exhaustively where the domain is small (all 2^24 colours for the HSV round trip, all
    byte pairs x a factor grid for blend), and against whole-pipeline outputs of the UNMODIFIED reference classes
    (tests/golden/aug_*.pt, written by oracle/make_golden.py --aug).

All uint8 work is bit-exact by construction; the final float32 normalisation uses IEEE single division / subtraction
exactly as torch's CPU kernels do, so the oracle's output equals the reference's bit for bit.
"""
import math
import random

import numpy as np

PRECISION_BITS = 32 - 8 - 2            # Resample.c: fixed-point coefficient scale for 8-bit images
OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_HUE = 0, 1, 2, 3
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------------------------------
# Pillow resampling
# ------------------------------------------------------------------------------------------------
def _bilinear_filter(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def resample_coeffs(in_size, out_size, in0=0.0, in1=None):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0).
    Returns (xmin [out], count [out], k [out, ksize] int64): out[x] = clip8((sum_j k[x,j] * in[xmin[x]+j] + 2^21) >> 22)."""
    if in1 is None:
        in1 = float(in_size)
    scale = (in1 - in0) / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = [_bilinear_filter((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        xmin[xx], cnt[xx] = lo, n
        for j, v in enumerate(w):
            kk[xx, j] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
    return xmin, cnt, kk


def nearest_coeffs(in_size, out_size):
    """Image.resize(NEAREST) (Geometry.c ImagingScaleAffine, nearest filter): source index = floor((x + 0.5) * scale);
    as a one-tap table of the same shape as resample_coeffs"""
    scale = in_size / out_size
    xmin = np.zeros(out_size, np.int32)
    xo = scale * 0.5                                   # a[2] + a[0] * 0.5, then += a[0] per pixel (double accumulation)
    for x in range(out_size):
        xi = int(math.floor(xo))
        xmin[x] = min(max(xi, 0), in_size - 1)
        xo += scale
    return xmin, np.ones(out_size, np.int32), np.full((out_size, 1), 1 << PRECISION_BITS, np.int64)


def identity_coeffs(size):
    return np.arange(size, dtype=np.int32), np.ones(size, np.int32), np.full((size, 1), 1 << PRECISION_BITS, np.int64)


def apply_tables(img, tx, ty):
    """img [H, W, C] uint8; tx / ty = (start, count, coeff, step) tables over the SOURCE coordinates: horizontal pass first
    (uint8 intermediate), then the vertical pass -- Resample.c ImagingResampleInner."""
    xs, xc, xk, xstep = tx
    ys, yc, yk, ystep = ty
    H, W, C = img.shape
    src = img.astype(np.int64)
    half = 1 << (PRECISION_BITS - 1)
    tmp = np.zeros((H, len(xs), C), np.int64)
    for x in range(len(xs)):
        acc = np.full((H, C), half, np.int64)
        for j in range(int(xc[x])):
            acc += src[:, int(xs[x]) + xstep * j, :] * int(xk[x, j])
        tmp[:, x, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.zeros((len(ys), len(xs), C), np.int64)
    for y in range(len(ys)):
        acc = np.full((len(xs), C), half, np.int64)
        for j in range(int(yc[y])):
            acc += tmp[int(ys[y]) + ystep * j] * int(yk[y, j])
        out[y] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# Pillow colour arithmetic
# ------------------------------------------------------------------------------------------------
def to_l(img):
    """convert("L"): ITU-R 601-2 luma, Convert.c L24 >> 16"""
    i = img.astype(np.int64)
    return ((i[..., 0] * 19595 + i[..., 1] * 38470 + i[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(a, b, alpha):
    """Image.blend(a, b, alpha) (Blend.c): float32 a + alpha * (b - a), truncated; clipped when extrapolating"""
    al = np.float32(alpha)
    d = (b.astype(np.int32) - a.astype(np.int32)).astype(np.float32)
    t = a.astype(np.float32) + al * d                     # float32 multiply, then float32 add (no fused multiply-add)
    if 0.0 <= float(al) <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    out = np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32)))
    return out.astype(np.uint8)


def rgb_to_hsv(img):
    """Convert.c rgb2hsv_row (after colorsys.py), uint8 in / out"""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    same = maxc == minc
    cr = np.where(same, 1, maxc - minc).astype(np.float32)
    mx = np.where(maxc == 0, 1, maxc).astype(np.float32)
    s = cr / mx
    rc = (maxc - r).astype(np.float32) / cr
    gc = (maxc - g).astype(np.float32) / cr
    bc = (maxc - b).astype(np.float32) / cr
    # h is a C float; the literals are doubles, so each expression is evaluated in double and rounded back on assignment
    h = np.where(r == maxc, bc - gc,
                 np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc.astype(np.float64)).astype(np.float32),
                          (4.0 + gc.astype(np.float64) - rc.astype(np.float64)).astype(np.float32)))
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    uh = np.where(same, 0, uh)
    us = np.where(same, 0, us)
    return np.stack([uh, us, maxc], -1).astype(np.uint8)


def hsv_to_rgb(img):
    """Convert.c hsv2rgb (after colorsys.py), uint8 in / out"""
    h, s, v = (img[..., i].astype(np.int64) for i in range(3))
    hf = h.astype(np.float32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i.astype(np.float32).astype(np.float64)).astype(np.float32)
    fs = (s.astype(np.float32).astype(np.float64) / 255.0).astype(np.float32)
    vf = v.astype(np.float32).astype(np.float64)
    f64, fs64 = f.astype(np.float64), fs.astype(np.float64)

    def rnd(x):                                           # C round(): half away from zero (all values are >= 0 here)
        return np.clip(np.floor(x + 0.5).astype(np.int64), 0, 255)
    p = rnd(vf * (1.0 - fs64))
    q = rnd(vf * (1.0 - fs64 * f64))
    t = rnd(vf * (1.0 - fs64 * (1.0 - f64)))
    sel = i % 6
    r = np.choose(sel, [v, q, p, p, t, v])
    g = np.choose(sel, [t, v, v, q, p, p])
    b = np.choose(sel, [p, p, t, v, v, q])
    grey = s == 0
    r, g, b = (np.where(grey, v, c) for c in (r, g, b))
    return np.stack([r, g, b], -1).astype(np.uint8)


def hue_shift_byte(hue_factor):
    """torchvision _functional_pil.adjust_hue: np.int32(hue_factor * 255).astype(np.uint8)"""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


def adjust(img, op, factor):
    """one ColorJitter step on an RGB uint8 image [H, W, 3]"""
    if op == OP_BRIGHTNESS:                               # ImageEnhance.Brightness: blend(black, img, f)
        return blend(np.zeros_like(img), img, factor)
    if op == OP_CONTRAST:                                 # ImageEnhance.Contrast: blend(mean grey, img, f)
        L = to_l(img)
        mean = int(float(L.astype(np.int64).sum()) / L.size + 0.5)      # ImageStat: sum / count in Python floats
        return blend(np.full_like(img, mean), img, factor)
    if op == OP_SATURATION:                               # ImageEnhance.Color: blend(L as RGB, img, f)
        L = to_l(img)
        return blend(np.repeat(L[..., None], 3, -1), img, factor)
    if op == OP_HUE:
        hsv = rgb_to_hsv(img)
        hsv[..., 0] = (hsv[..., 0].astype(np.int64) + hue_shift_byte(factor)).astype(np.uint8)     # uint8 wrap-around
        return hsv_to_rgb(hsv)
    raise ValueError(op)


def normalize(img, mean=MEAN, std=STD):
    """ToTensor + Normalize: [H, W, 3] uint8 -> [3, H, W] float32, IEEE single precision throughout"""
    x = img.astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32)
    s = np.asarray(std, np.float32)
    return np.ascontiguousarray(((x - m) / s).transpose(2, 0, 1))


# ------------------------------------------------------------------------------------------------
# parameter sampling: the reference's classes, call for call on Python's `random` and numpy's global generator
# ------------------------------------------------------------------------------------------------
class ClipPlan:
    """everything random about one clip: geometry tables (per clip) + per-frame grey / jitter decisions"""

    def __init__(self, n_frames, W, H):
        self.n_frames, self.W, self.H = n_frames, W, H
        self.box = (0, 0, W, H)          # crop box in source coordinates (x, y, w, h)
        self.flip_src = False            # flip applied BEFORE the crop / resize (acts on the source image)
        self.flip_out = False            # flip applied after the resize
        self.resize = None               # None | ('bilinear', (Wo, Ho)) | ('nearest', (Wo, Ho))
        self.out_crop = None             # None | (x, y, w, h): window of the RESIZED grid that is kept (CenterCrop after Scale)
        self.gray = [-1] * n_frames      # channel replicated into all three, or -1
        self.jitter = [[] for _ in range(n_frames)]      # per frame: [(op, factor), ...] in application order

    def out_size(self):
        if self.out_crop is not None:
            return self.out_crop[2], self.out_crop[3]
        return (self.box[2], self.box[3]) if self.resize is None else self.resize[1]

    def tables(self):
        """((start, count, coeff, step) for x, same for y) in SOURCE pixel coordinates"""
        x0, y0, w, h = self.box
        Wo, Ho = (w, h) if self.resize is None else self.resize[1]
        if self.resize is None:
            tx, ty = identity_coeffs(w), identity_coeffs(h)
        elif self.resize[0] == 'bilinear':
            tx, ty = resample_coeffs(w, Wo), resample_coeffs(h, Ho)
        else:
            tx, ty = nearest_coeffs(w, Wo), nearest_coeffs(h, Ho)
        xs, xc, xk = tx
        ys, yc, yk = ty
        if self.out_crop is not None:
            ox, oy, ow, oh = self.out_crop
            xs, xc, xk = xs[ox:ox + ow], xc[ox:ox + ow], xk[ox:ox + ow]
            ys, yc, yk = ys[oy:oy + oh], yc[oy:oy + oh], yk[oy:oy + oh]
        if self.flip_out:                                 # out[x] = resized[Wo - 1 - x]
            xs, xc, xk = xs[::-1].copy(), xc[::-1].copy(), xk[::-1].copy()
        if self.flip_src:                                 # crop coordinate u of the flipped source = column W - 1 - (x0 + u)
            return (self.W - 1 - (x0 + xs), xc, xk, -1), (y0 + ys, yc, yk, 1)
        return (x0 + xs, xc, xk, 1), (y0 + ys, yc, yk, 1)


def plan_random_sized_crop(plan, size, p=1.0, consistent=True):
    """augmentation.py:147-203 (consistent=True branch)"""
    assert consistent
    W, H = plan.box[2], plan.box[3]
    assert plan.box[:2] == (0, 0) and plan.resize is None
    if random.random() < p:
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
                plan.box = (x1, y1, w, h)
                plan.resize = ('bilinear', (size, size))
                return
        # fallback (:196-199): Scale(size, BILINEAR) -- short side to `size` (:27-38) -- then CenterCrop(size) (:51-58)
        if (W <= H and W == size) or (H <= W and H == size):
            ow, oh = W, H
        elif W < H:
            ow, oh = size, int(size * H / W)
            plan.resize = ('bilinear', (ow, oh))
        else:
            ow, oh = int(size * W / H), size
            plan.resize = ('bilinear', (ow, oh))
        x1 = int(round((ow - size) / 2.))
        y1 = int(round((oh - size) / 2.))
        if plan.resize is None:
            plan.box = (x1, y1, size, size)
        else:
            plan.out_crop = (x1, y1, size, size)
        return
    raise NotImplementedError('RandomSizedCrop with p < 1 (CenterCrop branch) is not restated')


def plan_random_crop(plan, size):
    """augmentation.py:98-144 (consistent=True, no flow map)"""
    W, H = plan.box[2], plan.box[3]
    th, tw = (size, size) if isinstance(size, int) else size
    if W == tw and H == th:
        return
    x1 = random.randint(0, W - tw)
    y1 = random.randint(0, H - th)
    plan.box = (plan.box[0] + x1, plan.box[1] + y1, tw, th)


def plan_scale(plan, size):
    """augmentation.py:20-41 with a (w, h) tuple and the default NEAREST interpolation (main.py:119)"""
    plan.resize = ('nearest', tuple(size))


def plan_flip(plan, threshold=0.5):
    """augmentation.py:206-232 (consistent=True); before or after the resize depending on where the recipe puts it"""
    if random.random() < threshold:
        if plan.resize is None and plan.box == (0, 0, plan.W, plan.H):
            plan.flip_src = not plan.flip_src
        else:
            plan.flip_out = not plan.flip_out


def plan_gray(plan, p=0.5):
    """augmentation.py:235-261 (consistent=False): per frame random.random(), then np.random.choice(3)"""
    for f in range(plan.n_frames):
        if random.random() < p:
            plan.gray[f] = int(np.random.choice(3))


def plan_jitter(plan, brightness=0.5, contrast=0.5, saturation=0.5, hue=0.25, p=1.0):
    """augmentation.py:264-355 (consistent=False): one random.random(), then per frame four uniforms and a shuffle"""
    b = [max(1 - brightness, 0), 1 + brightness]
    c = [max(1 - contrast, 0), 1 + contrast]
    s = [max(1 - saturation, 0), 1 + saturation]
    h = [-hue, hue]
    if random.random() < p:
        for f in range(plan.n_frames):
            ops = [(OP_BRIGHTNESS, random.uniform(b[0], b[1])), (OP_CONTRAST, random.uniform(c[0], c[1])),
                   (OP_SATURATION, random.uniform(s[0], s[1])), (OP_HUE, random.uniform(h[0], h[1]))]
            random.shuffle(ops)
            plan.jitter[f] = ops


def plan_k400(n_frames, W, H, img_dim):
    """main.py:125-133"""
    plan = ClipPlan(n_frames, W, H)
    plan_random_sized_crop(plan, img_dim)
    plan_flip(plan)
    plan_gray(plan)
    plan_jitter(plan)
    return plan


def plan_ucf101(n_frames, W, H, img_dim):
    """main.py:115-124"""
    plan = ClipPlan(n_frames, W, H)
    plan_flip(plan)
    plan_random_crop(plan, 224)
    plan_scale(plan, (img_dim, img_dim))
    plan_gray(plan)
    plan_jitter(plan)
    return plan


# ------------------------------------------------------------------------------------------------
def augment_clip(frames, plan, num_seq, seq_len):
    """frames [F, H, W, 3] uint8 -> (block [num_seq, 3, seq_len, Ho, Wo] float32, uint8 frames after the jitter)
    (dataset_3d.py:108-112: stack, view(num_seq, seq_len, C, H, W), transpose(1, 2))"""
    F = frames.shape[0]
    assert F == num_seq * seq_len == plan.n_frames
    tx, ty = plan.tables()
    u8, out = [], []
    for f in range(F):
        img = apply_tables(frames[f], tx, ty)
        if plan.gray[f] >= 0:
            img = np.repeat(img[..., plan.gray[f]:plan.gray[f] + 1], 3, -1)
        for op, factor in plan.jitter[f]:
            img = adjust(img, op, factor)
        u8.append(img)
        out.append(normalize(img))
    t = np.stack(out, 0)
    C, Ho, Wo = t.shape[1:]
    block = t.reshape(num_seq, seq_len, C, Ho, Wo).transpose(0, 2, 1, 3, 4)
    return np.ascontiguousarray(block), np.stack(u8, 0)


# ------------------------------------------------------------------------------------------------
# synthetic decoded frames + fingerprints shared by the fixture generator and the tests
# ------------------------------------------------------------------------------------------------
def make_frames(seed, F, H, W):
    """[F, H, W, 3] uint8: moving smooth colour fields + texture + a few saturated / grey regions (exercise every HSV branch)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    out = np.zeros((F, H, W, 3), np.uint8)
    ph = rng.uniform(0, 6.28, (3, 3))
    fr = rng.uniform(0.02, 0.09, (3, 2))
    for f in range(F):
        img = np.zeros((H, W, 3))
        for c in range(3):
            img[..., c] = 127 + 100 * np.sin(fr[c, 0] * xx + ph[c, 0] + 0.3 * f) * np.cos(fr[c, 1] * yy + ph[c, 1] - 0.2 * f)
        img += rng.normal(0, 12, (H, W, 3))
        img = np.clip(img, 0, 255)
        img[: H // 8, : W // 6] = 255                     # white
        img[-H // 8:, -W // 6:] = 0                       # black
        img[H // 3: H // 3 + H // 10, : W // 5] = img[H // 3: H // 3 + H // 10, : W // 5, :1]     # grey (s = 0)
        out[f] = img.astype(np.uint8)
    return out


def fingerprint(block):
    """sha256 of the float32 bytes + a strided sample (for diagnosis when the hash differs)"""
    import hashlib
    b = np.ascontiguousarray(block, dtype=np.float32)
    flat = b.reshape(-1)
    step = max(1, flat.size // 4096)
    return dict(shape=tuple(b.shape), sha256=hashlib.sha256(b.tobytes()).hexdigest(), step=step, sample=flat[::step][:4096].copy())
