"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU restatement, in numpy, of the per-sample input path the
device input kernel (csrc/input.cu) replaces:

  * MixDataset.__getitem__ (mix_dataset.py:37-90): partner draw, mixup (PIL `Image.blend` on uint8 images `:66`, or
    `img1.mul(ratio).add_(img2, alpha=1 - ratio)` on tensors `:63`), cutmix (`rand_bbox` `:10-24`, paste `:79,84`) and the
    recomputed ratio `:80,85`;
  * ToTensor + Normalize (factory.py:163-165,173-174; torchvision: `img.float().div(255)`, `sub_(mean).div_(std)`);
  * RandomErasing._erase (transforms.py:377-407; factory.py:178-182 uses mode="pixel", max_count=1): box draws, then the
    box is overwritten with N(0, 1) noise ("pixel") or zeros ("const").

The random DECISIONS are drawn from Python's `random` in exactly the reference's order (so the same seed gives the same
partner, ratio and boxes as the reference classes: pinned by tests/golden/input_ops.pt, written by
oracle/make_input_golden.py from the reference's own MixDataset / RandomErasing, torchvision and PIL).  The erase NOISE
cannot follow torch's CPU generator on a GPU; it is iid N(0, 1) from Philox4x32-10 (Salmon et al., SC'11; Random123 known
answers checked in tests) + Box-Muller, a pure function of (seed, y, x, channel) restated here bit-for-bit.
"""
import math

import numpy as np

F = np.float32
MEAN = (0.485, 0.456, 0.406)  # factory.py:163-165
STD = (0.229, 0.224, 0.225)
MODE_NONE, MODE_MIXUP, MODE_CUTMIX = 0, 1, 2
DOMAIN_U8, DOMAIN_F32 = 0, 1
ERASE_CONST, ERASE_PIXEL = 0, 1
TABLE_COLS = 24
PHILOX_KEY1 = 0x7674B200


# ------------------------------------------------------------------------------------------------ host decisions
def rand_bbox(rng, size, ratio):
    """mix_dataset.py:10-24.  size = (w, h) for PIL images, = tensor.shape[1:] = (h, w) for tensors (the reference passes
    both; the first entry bounds x)."""
    w, h = size
    r = math.sqrt(1 - ratio)
    cut_w, cut_h = int(w * r), int(h * r)
    cx = rng.randrange(w)
    cy = rng.randrange(h)
    clamp = lambda v, hi: min(max(v, 0), hi)  # noqa: E731
    return (clamp(cx - cut_w // 2, w), clamp(cy - cut_h // 2, h), clamp(cx + cut_w // 2, w), clamp(cy + cut_h // 2, h))


def draw_partner(rng, index, n):
    """mix_dataset.py:46-49."""
    j = index
    while j == index:
        j = rng.randrange(n)
    return j


def draw_mix(rng, index, size, mixup, cutmix):
    """mix_dataset.py:55-85 after the partner draw -> (mode, weight, box, ratio).  `weight` is the mixup ratio
    (the share of img1), `ratio` what __getitem__ returns (`:90`)."""
    apply_mixup, apply_cutmix = mixup > 0, cutmix > 0
    if apply_mixup and apply_cutmix:
        if index % 2 == 0:
            apply_cutmix = False
        else:
            apply_mixup = False
    if apply_mixup:
        ratio = rng.betavariate(mixup, mixup)
        return MODE_MIXUP, ratio, (0, 0, 0, 0), ratio
    if apply_cutmix:
        ratio = rng.uniform(0, 1) if cutmix == 1 else rng.betavariate(cutmix, cutmix)
        x1, y1, x2, y2 = rand_bbox(rng, size, ratio)
        return MODE_CUTMIX, 1.0, (x1, y1, x2, y2), 1 - ((x2 - x1) * (y2 - y1) / (size[0] * size[1]))
    return MODE_NONE, 1.0, (0, 0, 0, 0), 1


def draw_erase(rng, img_h, img_w, p, min_area=0.02, max_area=1 / 3, min_aspect=0.3, max_aspect=None, min_count=1,
               max_count=None):
    """transforms.py:377-407 -> list of (top, left, h, w) boxes (empty when the coin flip or all 10 attempts fail)."""
    max_aspect = max_aspect or 1 / min_aspect
    log_ar = (math.log(min_aspect), math.log(max_aspect))
    max_count = max_count or min_count
    if rng.random() > p:
        return []
    area = img_h * img_w
    count = min_count if min_count == max_count else rng.randint(min_count, max_count)
    boxes = []
    for _ in range(count):
        for _attempt in range(10):
            target = rng.uniform(min_area, max_area) * area / count
            ar = math.exp(rng.uniform(*log_ar))
            h = int(round(math.sqrt(target * ar)))
            w = int(round(math.sqrt(target / ar)))
            if w < img_w and h < img_h:
                top = rng.randint(0, img_h - h)
                left = rng.randint(0, img_w - w)
                boxes.append((top, left, h, w))
                break
    return boxes


# ------------------------------------------------------------------------------------------------ arithmetic
def normalize_lut(mean=MEAN, std=STD):
    """[3, 256] float32: ((v / 255) - mean_c) / std_c in float32 steps (torchvision to_tensor + normalize)."""
    v = np.arange(256, dtype=F) / F(255)
    return np.stack([((v - F(m)) / F(s)).astype(F) for m, s in zip(mean, std)])


def pil_blend(a, b, alpha):
    """PIL ImagingBlend for 0 <= alpha <= 1 (what `Image.blend(img1, img2, alpha)` runs, mix_dataset.py:66):
    out = (uint8)(in1 + alpha * (in2 - in1)) in float32, truncated."""
    al = F(alpha)
    d = (b.astype(np.int32) - a.astype(np.int32)).astype(F)
    return (a.astype(F) + al * d).astype(F).astype(np.int32).astype(np.uint8)


def _mulhilo(a, b):
    p = a.astype(np.uint64) * np.uint64(b)
    return (p >> np.uint64(32)).astype(np.uint32), (p & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Random123).  Counter words as uint32 arrays (broadcastable), key words as ints."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)])
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        hi0, lo0 = _mulhilo(c0, 0xD2511F53)
        hi1, lo1 = _mulhilo(c2, 0xCD9E8D57)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _uniform(x):
    """uint32 -> float32 in (0, 1): the top 24 bits, centred."""
    return ((x >> np.uint32(8)).astype(F) * F(2.0 ** -24) + F(2.0 ** -25)).astype(F)


def erase_noise(seed, ys, xs):
    """N(0, 1) noise of the pixels (ys, xs) of an erased box -> float32 [3, ...]: one Philox call per pixel
    (counter (x, y, 0, 0), key (seed, PHILOX_KEY1)), Box-Muller on words (0, 1) -> channels 0 and 1, on words (2, 3) ->
    channel 2 (the sine half is dropped)."""
    r0, r1, r2, r3 = philox4x32_10(np.asarray(xs, np.uint32), np.asarray(ys, np.uint32), 0, 0, seed, PHILOX_KEY1)
    two_pi = F(6.2831853071795864769)
    ra = np.sqrt(F(-2.0) * np.log(_uniform(r0))).astype(F)
    ta = (two_pi * _uniform(r1)).astype(F)
    rb = np.sqrt(F(-2.0) * np.log(_uniform(r2))).astype(F)
    tb = (two_pi * _uniform(r3)).astype(F)
    return np.stack([ra * np.cos(ta), ra * np.sin(ta), rb * np.cos(tb)]).astype(F)


def _to_chw(u8_hwc, lut):
    return np.stack([lut[c][u8_hwc[..., c]] for c in range(3)])


def _erase(chw, box, mode, seed):
    top, left, h, w = box
    if h <= 0 or w <= 0:
        return chw
    out = chw.copy()
    if mode == ERASE_PIXEL:
        ys, xs = np.meshgrid(np.arange(top, top + h), np.arange(left, left + w), indexing="ij")
        out[:, top:top + h, left:left + w] = erase_noise(seed, ys, xs)
    else:
        out[:, top:top + h, left:left + w] = 0
    return out


def input_batch(src, table, mean=MEAN, std=STD):
    """The whole device op.  src uint8 [S, H, W, 3]; table int32 [B, TABLE_COLS] (layout: include/vtb200.h, written by
    device_input.pack_table) -> float32 [B, 3, H, W]."""
    lut = normalize_lut(mean, std)
    S, H, W, _ = src.shape
    out = np.empty((table.shape[0], 3, H, W), F)
    fbits = lambda v: np.array([v], np.int32).view(F)[0]  # noqa: E731
    for b, row in enumerate(np.asarray(table, np.int32)):
        s1, s2, mode, domain = (int(v) for v in row[:4])
        w1, w2 = fbits(row[4]), fbits(row[5])
        x1, y1, x2, y2 = (int(v) for v in row[6:10])
        box_a, box_b = tuple(int(v) for v in row[10:14]), tuple(int(v) for v in row[14:18])
        seed_a, seed_b, emode = int(row[18]) & 0xFFFFFFFF, int(row[19]) & 0xFFFFFFFF, int(row[20])
        a, p = src[s1], src[s2]
        if domain == DOMAIN_U8:
            if mode == MODE_MIXUP:
                a = pil_blend(a, p, w1)
            elif mode == MODE_CUTMIX:
                a = a.copy()
                a[y1:y2, x1:x2] = p[y1:y2, x1:x2]
            out[b] = _erase(_to_chw(a, lut), box_a, emode, seed_a)
        else:
            va = _erase(_to_chw(a, lut), box_a, emode, seed_a)
            if mode == MODE_NONE:
                out[b] = va
                continue
            vb = _erase(_to_chw(p, lut), box_b, emode, seed_b)
            if mode == MODE_MIXUP:
                t = (va * w1).astype(F)  # img1.mul(ratio)
                out[b] = (t.astype(np.float64) + vb.astype(np.float64) * np.float64(w2)).astype(F)  # add_(img2, alpha): fma
            else:
                va = va.copy()
                va[:, y1:y2, x1:x2] = vb[:, y1:y2, x1:x2]
                out[b] = va
    return out
