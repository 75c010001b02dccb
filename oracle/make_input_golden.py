"""Generates tests/golden/input_ops.pt by running the REFERENCE's own loader classes — mix_dataset.MixDataset /
rand_bbox and transforms.RandomErasing — with torchvision's ToTensor / Normalize and PIL's Image.blend / paste (the
functions factory.py:159-226 composes) over small seeded uint8 datasets, in both orders the reference supports
(mix_before_aug = True: mix PIL images, then ToTensor / Normalize / RandomErasing;  False: the full transform per image,
then mix tensors).  RandomErasing runs in mode "const" here so the fixtures are deterministic (mode "pixel" draws torch
CPU noise, which no device kernel can follow; its box geometry is the same code path).
Build container only (needs /root/reference):  python oracle/make_input_golden.py
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import hashlib
import os
import random
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms as T

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]  # factory.py:163-165

CASES = [
    # name, H, W, n, mixup, cutmix, erasing, mix_before_aug, seed
    ("pil_mix_cut_erase", 20, 28, 8, 0.8, 1.0, 0.6, True, 11),
    ("pil_cut_beta_odd", 18, 15, 6, 0.0, 0.5, 0.0, True, 12),
    ("pil_mixup_only", 16, 16, 5, 0.2, 0.0, 0.0, True, 13),
    ("tensor_mix_cut_erase", 20, 28, 8, 0.8, 1.0, 0.6, False, 14),
    ("tensor_mixup_only", 16, 16, 5, 0.2, 0.0, 0.0, False, 15),
    ("tensor_cut_odd_erase", 18, 15, 6, 0.0, 1.0, 0.7, False, 16),
    ("plain_erase", 24, 20, 6, 0.0, 0.0, 0.5, True, 17),
]


class Images:
    """Stands in for LMDBDataset (dataset.py): `[i] -> (transform(PIL image), label)`."""

    def __init__(self, u8, transform):
        self.u8, self.transform = u8, transform

    def __len__(self):
        return len(self.u8)

    def __getitem__(self, i):
        return self.transform(Image.fromarray(self.u8[i])), int(i) * 7 % 5


def main():
    ref_mix = ref_loader.load_reference_module("mix_dataset")
    ref_tf = ref_loader.load_reference_module("transforms")
    out = {"mean": MEAN, "std": STD, "cases": {}}
    for name, H, W, n, mixup, cutmix, erasing, before, seed in CASES:
        u8 = np.random.default_rng(seed).integers(0, 256, (n, H, W, 3), dtype=np.uint8)
        tail = [T.ToTensor(), T.Normalize(mean=MEAN, std=STD)]
        if erasing > 0:  # factory.py:176-182 (mode "const" instead of "pixel": see the header)
            tail.append(ref_tf.RandomErasing(erasing, mode="const", max_count=1, num_splits=0, device="cpu"))
        if before:  # factory.py:184-190 with an empty pre-transform (crop / flip / RandAugment stay on the host)
            ds = ref_mix.MixDataset(Images(u8, lambda im: im), T.Compose(tail), mixup, cutmix)
        else:
            ds = ref_mix.MixDataset(Images(u8, T.Compose(tail)), T.Compose([]), mixup, cutmix)
        random.seed(seed)
        items = [ds[i] for i in range(n)]
        out["cases"][name] = {
            "H": H, "W": W, "n": n, "mixup": mixup, "cutmix": cutmix, "erasing": erasing, "mix_before_aug": before,
            "seed": seed, "u8": torch.from_numpy(u8), "img": torch.stack([it[0] for it in items]),
            "label1": [it[1] for it in items], "label2": [it[2] for it in items], "ratio": [float(it[3]) for it in items],
        }
    # PIL blend known answers: every (a, b) byte pair at a few alphas, pinned by SHA-256 of the [256, 256] result
    # (a = row index, b = column index, mode "L")
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    alphas = [0.0, 1.0, 0.5, 0.26839999999999997, 0.123456789, 0.9999, 0.731]
    out["blend"] = {"alphas": alphas, "sha256": [
        hashlib.sha256(np.asarray(Image.blend(Image.fromarray(a), Image.fromarray(a.T.copy()), al)).tobytes()).hexdigest()
        for al in alphas]}
    # rand_bbox draws
    random.seed(5)
    out["bbox"] = [(size, r, ref_mix.rand_bbox(size, r)) for size, r in
                   [((28, 20), 0.3), ((20, 28), 0.9), ((224, 224), 0.5), ((15, 18), 0.01), ((7, 7), 0.999)]]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_ops.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
