"""Micro-benchmark of the device input kernel (vtb_input_batch) at the BASELINE batch (256 x 224 x 224):
device-resident uint8 sources (kernel alone, against the HBM roofline) and from pinned host memory (H2D inside)."""
import json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import device_input as D

B, H, W = 256, 224, 224
peak = 6543.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
pipe = D.DeviceInput()
src_host = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8).pin_memory()
src_dev = src_host.cuda()
out = torch.empty(B, 3, H, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2
same = {i: i for i in range(B)}
CASES = {"none (normalise only)": dict(mixup=0.0, cutmix=0.0, erasing=0.0),
         "swin-s conf (mixup 0.8 / cutmix 1.0 / erasing 0.25)": dict(mixup=0.8, cutmix=1.0, erasing=0.25),
         "tensor order, erasing 0.25": dict(mixup=0.8, cutmix=1.0, erasing=0.25, mix_before_aug=False)}
for name, kw in CASES.items():
    s = D.MixSampler(rng=random.Random(0), **kw)
    ds = [s.sample(i, B, H, W) for i in range(B)]
    table = D.pack_table(ds, same, s.mix_before_aug, "pixel")
    mixed = sum(d.mode != 0 for d in ds)
    nbytes = B * H * W * 12 + (B + mixed) * H * W * 3  # algorithmic: out written once, each source byte read once per use
    for label, src in (("device", src_dev), ("pinned host", src_host)):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            pipe(src, table, out=out)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        print(f"{name:52s} {label:12s} {ms*1e3:8.1f} us  {B/ms*1e3:10.0f} img/s  {nbytes/ms/1e6:7.0f} GB/s "
              f"({nbytes/ms/1e6/peak*100:4.1f} % of {peak:.0f})", flush=True)
