import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/vision-transformers-pytorch_b200")
import torch
import bench
B = int(os.environ.get("B", 128))
m = bench.build_model("pvt_small").cuda().train()
x = torch.randn(B, 3, 224, 224, device="cuda")
for i in range(3):
    out = m(x)
    torch.cuda.synchronize(); print("fwd ok", i, flush=True)
    out.float().sum().backward()
    torch.cuda.synchronize(); print("bwd ok", i, flush=True)
