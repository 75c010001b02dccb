"""Device input path: the per-sample work of the reference's loader — `MixDataset.__getitem__`
(mix_dataset.py:37-90), `ToTensor` + `Normalize` (factory.py:163-174) and `RandomErasing` (transforms.py:321-418,
factory.py:178-182) — split into

  * `MixSampler`: the random DECISIONS (partner index, mixup / cutmix ratio, cutmix box, erase boxes), drawn on the host
    from a `random.Random` in exactly the reference's order, so one seed gives the reference's partner / ratio / boxes;
  * `DeviceInput`: ONE kernel launch (`vtb_input_batch`, csrc/input.cu) that turns the uint8 HWC images into the
    normalised fp32 NCHW batch with those decisions applied.  The host ships uint8 (4x fewer PCIe bytes than the
    reference's fp32 batches, train.py:267) and does no per-pixel arithmetic.

What stays on the host: decode, RandomResizedCrop, flip and RandAugment (PIL ops, factory.py:168-171).  With
`mix_before_aug=True` AND RandAugment active the reference mixes BEFORE RandAugment (factory.py:184-186); that order
cannot be split across the PCIe bus, so such a loader keeps the PIL mix on the host and uses mode "none" rows here
(normalise + erase on the device).  There is no CPU fallback: `DeviceInput` raises without CUDA / the built library.
"""
import math
import random

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)  # factory.py:163-165
STD = (0.229, 0.224, 0.225)
MODE_NONE, MODE_MIXUP, MODE_CUTMIX = 0, 1, 2
DOMAIN_U8, DOMAIN_F32 = 0, 1
ERASE_MODES = {"const": 0, "pixel": 1}
TABLE_COLS = 24


def rand_bbox(size, ratio, rng=random):
    """Same contract as mix_dataset.rand_bbox (`:10-24`): `size` = (w, h) of a PIL image — or `tensor.shape[1:]`, which
    the reference passes for tensors — and the cut covers a `1 - ratio` share of it around a uniformly drawn centre."""
    w, h = size
    side = math.sqrt(1 - ratio)
    half_w, half_h = int(w * side) // 2, int(h * side) // 2
    cx = rng.randrange(w)
    cy = rng.randrange(h)
    return (max(cx - half_w, 0), max(cy - half_h, 0), min(cx + half_w, w), min(cy + half_h, h))


class Decision:
    """One output sample: which sources, how they are mixed, what is erased; `ratio` is what the loader yields."""
    __slots__ = ("index", "partner", "mode", "weight", "box", "ratio", "erase_a", "erase_b", "seed_a", "seed_b")

    def __init__(self, index):
        self.index, self.partner, self.mode, self.weight, self.box, self.ratio = index, index, MODE_NONE, 1.0, (0,) * 4, 1
        self.erase_a = self.erase_b = (0, 0, 0, 0)
        self.seed_a = self.seed_b = 0


class MixSampler:
    """Draws `Decision`s.  Arguments mirror MixDataset(mixup, cutmix) (mix_dataset.py:28-32), `erasing` / `mix_before_aug`
    of make_dataset (factory.py:159-186) and RandomErasing's box parameters (transforms.py:341-353; the reference uses
    max_count = 1, so at most one box per source).  `rng` is consumed in the reference's order:
      mix_before_aug=True : partner, mix draws, erase draws of the result          (erase is part of the post-transform)
      mix_before_aug=False: erase draws of img1, partner, erase draws of img2, mix draws   (erase is part of dataset[i])
    """

    def __init__(self, mixup=0.2, cutmix=1, erasing=0.0, mix_before_aug=True, min_area=0.02, max_area=1 / 3,
                 min_aspect=0.3, max_aspect=None, rng=None, noise_seed=0):
        self.mixup, self.cutmix, self.erasing, self.mix_before_aug = mixup, cutmix, erasing, mix_before_aug
        self.min_area, self.max_area = min_area, max_area
        self.log_aspect = (math.log(min_aspect), math.log(max_aspect or 1 / min_aspect))
        self.rng = rng if rng is not None else random
        self._noise = random.Random(noise_seed)  # separate stream: keeps `rng` in step with the reference

    def _erase_box(self, H, W):
        """transforms.py:377-407 with min_count = max_count = 1 -> (top, left, h, w), h = 0 when nothing is erased."""
        rng = self.rng
        if self.erasing <= 0 or rng.random() > self.erasing:
            return (0, 0, 0, 0)
        for _ in range(10):
            target = rng.uniform(self.min_area, self.max_area) * (H * W)
            aspect = math.exp(rng.uniform(*self.log_aspect))
            h, w = int(round(math.sqrt(target * aspect))), int(round(math.sqrt(target / aspect)))
            if w < W and h < H:
                top = rng.randint(0, H - h)
                return (top, rng.randint(0, W - w), h, w)
        return (0, 0, 0, 0)

    def sample(self, index, n_dataset, H, W):
        d = Decision(index)
        rng, tensor_order = self.rng, not self.mix_before_aug
        use_mixup, use_cutmix = self.mixup > 0, self.cutmix > 0
        mixing = use_mixup or use_cutmix
        if tensor_order and self.erasing > 0:
            d.erase_a = self._erase_box(H, W)
        if mixing:
            while d.partner == index:
                d.partner = rng.randrange(n_dataset)
            if tensor_order and self.erasing > 0:
                d.erase_b = self._erase_box(H, W)
            if use_mixup and use_cutmix:  # alternate by dataset index parity (mix_dataset.py:55-60)
                use_mixup, use_cutmix = index % 2 == 0, index % 2 != 0
            if use_mixup:
                d.mode, d.weight = MODE_MIXUP, rng.betavariate(self.mixup, self.mixup)
                d.ratio = d.weight
            else:
                lam = rng.uniform(0, 1) if self.cutmix == 1 else rng.betavariate(self.cutmix, self.cutmix)
                size = (H, W) if tensor_order else (W, H)
                x1, y1, x2, y2 = d.box = rand_bbox(size, lam, rng)
                d.mode, d.ratio = MODE_CUTMIX, 1 - ((x2 - x1) * (y2 - y1) / (H * W))
        if not tensor_order and self.erasing > 0:
            d.erase_a = self._erase_box(H, W)
        d.seed_a, d.seed_b = self._noise.getrandbits(32), self._noise.getrandbits(32)
        return d


def _f32_bits(v):
    return int(np.array([v], np.float32).view(np.int32)[0])


def _as_i32(u):
    return u - (1 << 32) if u >= (1 << 31) else u


def pack_table(decisions, src_slots, mix_before_aug=True, erase_mode="pixel"):
    """int32 [B, 24] table of include/vtb200.h (vtb_input_batch).  `src_slots[dataset_index]` = row of the uint8 source
    batch holding that image."""
    if erase_mode not in ERASE_MODES:
        raise ValueError(f"erase_mode must be one of {sorted(ERASE_MODES)} ('rand' is not used by the reference's loader)")
    t = np.zeros((len(decisions), TABLE_COLS), np.int32)
    for row, d in zip(t, decisions):
        row[0], row[1], row[2] = src_slots[d.index], src_slots[d.partner], d.mode
        row[3] = DOMAIN_U8 if mix_before_aug else DOMAIN_F32
        if d.mode == MODE_MIXUP:
            # PIL: Image.blend(img1, img2, 1 - ratio) (`:66`);  tensors: img1.mul(ratio).add_(img2, alpha=1 - ratio) (`:63`)
            row[4] = _f32_bits(1 - d.weight) if mix_before_aug else _f32_bits(d.weight)
            row[5] = _f32_bits(1 - d.weight)
        row[6:10] = d.box
        row[10:14], row[14:18] = d.erase_a, d.erase_b
        row[18], row[19], row[20] = _as_i32(d.seed_a), _as_i32(d.seed_b), ERASE_MODES[erase_mode]
    return t


def check_table(table, n_src, H, W):
    """Host-side validation of a table (the kernel trusts it): raises ValueError on anything that would read or write
    outside the batch."""
    t = np.asarray(table)
    if t.dtype != np.int32 or t.ndim != 2 or t.shape[1] != TABLE_COLS:
        raise ValueError(f"table must be int32 [B, {TABLE_COLS}], got {t.dtype} {t.shape}")
    if t.size == 0:
        return
    if (t[:, :2] < 0).any() or (t[:, :2] >= n_src).any():
        raise ValueError(f"table: source index outside [0, {n_src})")
    if (t[:, 2] < 0).any() or (t[:, 2] > 2).any() or (t[:, 3] < 0).any() or (t[:, 3] > 1).any() or \
            (t[:, 20] < 0).any() or (t[:, 20] > 1).any():
        raise ValueError("table: mode must be 0..2, domain and erase mode 0..1")
    for c0 in (10, 14):
        top, left, h, w = (t[:, c0 + k].astype(np.int64) for k in range(4))
        if (h < 0).any() or (w < 0).any() or (top < 0).any() or (left < 0).any() or (top + h > H).any() or \
                (left + w > W).any():
            raise ValueError("table: erase box outside the image")
    w1 = t[:, 4].copy().view(np.float32)
    if not np.isfinite(w1[t[:, 2] == MODE_MIXUP]).all() or ((t[:, 2] == MODE_MIXUP) & (t[:, 3] == DOMAIN_U8) &
                                                           ((w1 < 0) | (w1 > 1))).any():
        raise ValueError("table: mixup weight must be finite (and a blend alpha in [0, 1] in the uint8 domain)")


class DeviceInput:
    """`batch = DeviceInput()(src_u8, table)`: src_u8 uint8 [S, H, W, 3] (CPU, ideally pinned, or CUDA), table int32
    [B, 24] (numpy / CPU tensor) -> fp32 [B, 3, H, W] on the device, one `vtb_input_batch` launch on the current stream."""

    def __init__(self, mean=MEAN, std=STD, device="cuda"):
        from vtb200 import lib as _lib

        self._lib = _lib
        _lib.get()  # raises without CUDA or without the built library: there is no CPU path
        self.mean, self.std, self.device = tuple(float(v) for v in mean), tuple(float(v) for v in std), torch.device(device)
        if len(self.mean) != 3 or len(self.std) != 3:
            raise ValueError("mean / std must have three entries")

    def __call__(self, src_u8, table, out=None):
        import ctypes as C

        from vtb200 import ops

        if src_u8.dtype != torch.uint8 or src_u8.dim() != 4 or src_u8.shape[3] != 3:
            raise ValueError(f"src must be uint8 [S, H, W, 3], got {src_u8.dtype} {tuple(src_u8.shape)}")
        S, H, W, _ = src_u8.shape
        tab = table.numpy() if isinstance(table, torch.Tensor) else np.ascontiguousarray(table)
        check_table(tab, S, H, W)
        B = tab.shape[0]
        src_dev = src_u8.to(self.device, non_blocking=True).contiguous()
        tab_dev = torch.from_numpy(np.ascontiguousarray(tab)).to(self.device, non_blocking=True)
        if out is None:
            out = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        elif out.shape != (B, 3, H, W) or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
            raise ValueError("out must be a contiguous fp32 CUDA tensor [B, 3, H, W]")
        lib = self._lib.get()
        f3 = C.c_float * 3
        with torch.cuda.device(self.device):
            self._lib.check(lib.vtb_input_batch(C.c_void_p(src_dev.data_ptr()), S, C.c_void_p(tab_dev.data_ptr()), B, H, W,
                                                f3(*self.mean), f3(*self.std), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), lib)
        ops.LAUNCHES += 1
        src_dev.record_stream(torch.cuda.current_stream())
        tab_dev.record_stream(torch.cuda.current_stream())
        return out


def make_batch(indices, fetch, n_dataset, sampler, device_input, labels=None, erase_mode="pixel"):
    """The batch a reference DataLoader over MixDataset would yield (train.py:265: input, label1, label2, ratio), built
    on the device.  `fetch(i)` -> uint8 HWC array of dataset item i (already cropped / flipped by the host pipeline);
    partners are fetched once even when several samples share them."""
    first = np.asarray(fetch(indices[0]))
    H, W = first.shape[:2]
    decisions = [sampler.sample(i, n_dataset, H, W) for i in indices]
    slots, images = {}, []
    for i in [d.index for d in decisions] + [d.partner for d in decisions]:
        if i not in slots:
            slots[i] = len(images)
            images.append(first if i == indices[0] else np.asarray(fetch(i)))
    src = torch.from_numpy(np.stack(images))
    if torch.cuda.is_available():
        src = src.pin_memory()
    table = pack_table(decisions, slots, sampler.mix_before_aug, erase_mode)
    batch = device_input(src, table)
    ratio = torch.tensor([float(d.ratio) for d in decisions], dtype=torch.float64)
    if labels is None:
        return batch, decisions, ratio
    label1 = torch.tensor([labels[d.index] for d in decisions])
    label2 = torch.tensor([labels[d.partner] for d in decisions])
    return batch, label1, label2, ratio
