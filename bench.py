#!/usr/bin/env python
"""bench.py — images/sec, fwd+bwd, on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload vit_b16|swin_s|pvt_small|halo_t|twins_s|dino_deit_s] [--impl reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1); rank 0 prints ONE JSON line.
A step = forward + cross-entropy + backward of the whole model on one synthetic batch (256 images/GPU, weak
scaling), gradients of every parameter produced and (N > 1) all-reduced — the reference's step body
train.py:265-283 without the optimizer (metric: "images/sec fwd+bwd").  `value` has inputs resident in HBM;
`e2e` runs the same step from pinned HOST buffers (H2D of the batch + D2H of the loss inside the timed region).
--impl reference times the oracle port of the reference path on the host cores (oracle/restate.py).
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vision-transformers-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "images/sec fwd+bwd"
# fwd+bwd GFLOP per image: matmul+bmm+conv FLOPs of the reference graph (BASELINE.md §2 / SURVEY §8d)
WORKLOADS = {
    "vit_b16": dict(gflop=105.147, batch=256, desc="ViT-B/16 224x224 fwd+bwd, batch 256/GPU"),
    "swin_s": dict(gflop=52.416, batch=256, desc="Swin-S 224x224 fwd+bwd, batch 256/GPU"),
    "pvt_small": dict(gflop=22.875, batch=128, desc="PVT-Small 224x224 fwd+bwd, batch 128/GPU"),
    "halo_t": dict(gflop=29.36, batch=128, desc="Halo-T* 224x224 fwd+bwd, batch 128/GPU"),
    "vit_tiny": dict(gflop=7.463, batch=64, desc="ViT-Tiny/16 224x224 fwd+bwd (plumbing)"),
    # not a BASELINE config: Twins-SVT with the paper's "S" widths through the reference's own ctor; FLOPs counted
    # with FlopCounterMode on the reference module like the others
    "twins_s": dict(gflop=34.486, batch=128, desc="Twins-SVT-S* 224x224 fwd+bwd, batch 128/GPU"),
    # BASELINE config 5: per source image 2x224^2 + 8x96^2 student fwd+bwd, 2x224^2 teacher fwd, + head
    "dino_deit_s": dict(gflop=113.4, batch=128, desc="DINO DeiT-S/16 multi-crop (2x224^2 + 8x96^2), 128 source images/GPU: "
                                                      "teacher fwd + student fwd/bwd + DINO loss + EMA"),
}


def build_model(workload, drop_path=None):
    import models

    if workload == "vit_b16":
        dp = 0.1 if drop_path is None else drop_path
        return models.VisionTransformer(models.FusedLinear(768, 1000), 224, 16, 12, 768, 12, 3072, 0., 0., 0., dp)
    if workload == "vit_tiny":
        return models.VisionTransformer(models.FusedLinear(192, 1000), 224, 16, 12, 192, 3, 768, 0., 0., 0., 0.)
    if workload == "swin_s":
        dp = 0.3 if drop_path is None else drop_path
        return models.SwinTransformer((224, 224), 1000, (2, 2, 18, 2), (96, 192, 384, 768), 32, (3, 6, 12, 24),
                                      (384, 768, 1536, 3072), 7, drop_path=dp)
    if workload == "pvt_small":
        return models.PyramidVisionTransformer(224, 1000, 3, (3, 4, 6, 3), (64, 128, 320, 512), (1, 2, 5, 8),
                                               (512, 1024, 1280, 2048), (8, 4, 2, 1), drop_path=0.1)
    if workload == "dino_deit_s":  # config/dino_deit-s-16.conf:1-19
        return models.dino(image_size=224, window_size=16, depth=12, dim=384, n_head=6, dim_ff=1536, dropout=0.,
                           drop_attn=0., drop_ff=0., drop_path=0.1 if drop_path is None else drop_path,
                           dim_head_out=65536, use_bn=False, norm_last_layer=False, depth_head=3, dim_head_ff=2048,
                           dim_head_bottleneck=256)
    if workload == "twins_s":
        import models.twins

        return models.twins.TwinsSVT(1000, (2, 2, 10, 4), (64, 128, 256, 512), 32, (2, 4, 8, 16), (256, 512, 1024, 2048), 7,
                                     drop_path=0.1 if drop_path is None else drop_path)
    if workload == "halo_t":
        return models.HaloTransformer((224, 224), 1000, (2, 2, 6, 2), (96, 192, 384, 768), 32, (3, 6, 12, 24),
                                      (384, 768, 1536, 3072), window_size=7, halo_size=3, drop_path=0.1)
    raise KeyError(workload)


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1410.6), d.get("hbm_gbs", 6452.5), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_step(workload, batch, threads, device="cpu", autocast=False):
    """One fwd+bwd of the oracle port (plain PyTorch ops) — on the host by default (`cpu_baseline`, `--impl reference`);
    with device="cuda", autocast=True it is the opt-in `--eager-baseline` leg: eager PyTorch under bf16 autocast on the same
    GPU, the library path the reference itself would take.  Returns a closure running one step."""
    import contextlib

    from oracle import restate as R
    import models

    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    model = build_model(workload, drop_path=0.0)
    sd = {k: (v.detach().clone().to(device).requires_grad_(v.is_floating_point())) for k, v in model.state_dict().items()}
    cast = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast else contextlib.nullcontext
    if workload == "dino_deit_s":
        xs = [torch.randn(batch, 3, 224, 224, device=device) for _ in range(2)] + \
             [torch.randn(batch, 3, 96, 96, device=device) for _ in range(8)]
        tsd = {k: v.detach().clone() for k, v in sd.items()}
        center = torch.zeros(1, 65536, device=device)

        def dino_step():
            for v in sd.values():
                if v.is_floating_point():
                    v.grad = None
            with cast():
                with torch.no_grad():
                    t_out = R.vit_forward(tsd, xs[:2], patch=16, depth=12, heads=6, head_fn=lambda f: R.dino_head(tsd, f))
                s_out = R.vit_forward(sd, xs, patch=16, depth=12, heads=6, head_fn=lambda f: R.dino_head(sd, f))
                loss = R.dino_loss(s_out.float(), t_out.float(), center, 10)
            loss.backward()
            with torch.no_grad():
                for k in tsd:
                    tsd[k].mul_(0.996).add_(sd[k].detach(), alpha=0.004)
            return loss.item()

        return dino_step
    x = torch.randn(batch, 3, 224, 224, device=device)
    y = torch.randint(0, 1000, (batch,), device=device)

    def fwd():
        if workload in ("vit_b16", "vit_tiny"):
            heads = 12 if workload == "vit_b16" else 3
            return R.vit_forward(sd, x, patch=16, depth=12, heads=heads,
                                 head_fn=lambda f: R.linear(f, sd["head.weight"], sd["head.bias"]))
        if workload == "swin_s":
            return R.swin_forward(sd, x, depths=(2, 2, 18, 2), n_heads=(3, 6, 12, 24), dim_head=32, window=7)
        if workload == "pvt_small":
            return R.pvt_forward(sd, x, depths=(3, 4, 6, 3), n_heads=(1, 2, 5, 8), reductions=(8, 4, 2, 1))
        if workload == "twins_s":
            return R.twins_forward(sd, x, depths=(2, 2, 10, 4), n_heads=(2, 4, 8, 16), dim_head=32, window=7)
        return R.halo_forward(sd, x, depths=(2, 2, 6, 2), n_heads=(3, 6, 12, 24), dim_head=32, window=7, halo=3)

    def step():
        for v in sd.values():
            if v.is_floating_point():
                v.grad = None
        with cast():
            loss = torch.nn.functional.cross_entropy(fwd().float(), y)
        loss.backward()
        return loss.item()

    return step


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 2 if args.workload == "dino_deit_s" else 8
    step = cpu_reference_step(args.workload, batch, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = batch / dt
    sample = f"{WORKLOADS[args.workload]['desc'].split(',')[0]} fp32, batch {batch} per step, {steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_step": batch, "where": "host CPU, oracle port of the reference path"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="vit_b16", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--impl", default="vtb200", choices=["vtb200", "reference"])
    ap.add_argument("--reducer", default="ddp", choices=["ddp", "flat"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dino-graph", action="store_true",
                    help="DINO at N > 1 (opt-in): capture the step as a CUDA graph with the centre update deferred — the column "
                         "sum of the teacher logits is written to a static buffer inside the graph, its all-reduce and the EMA "
                         "of the centre run right after the replay (same arithmetic: the loss of a step never sees its own "
                         "centre update, loss.py:119-152)")
    ap.add_argument("--eager-baseline", action="store_true",
                    help="extra leg (opt-in): the oracle port (plain PyTorch ops) in eager mode under bf16 autocast on the same "
                         "GPU — the library path the reference itself would take; a reported baseline")
    ap.add_argument("--e2e-u8", action="store_true",
                    help="extra leg (opt-in): the step fed through the device input path — pinned uint8 HWC batches + the "
                         "mixup / cutmix / erasing table -> vtb_input_batch -> step (SURVEY 8f rank 4)")
    ap.add_argument("--no-weight-arena", action="store_true",
                    help="cast every weight to bf16 in its own launch (autocast's behaviour) instead of one multi-tensor "
                         "cast per forward (vtb200.multi.enable_weight_arena)")
    ap.add_argument("--no-optimizer-leg", action="store_true",
                    help="skip the secondary `with_optimizer` measurement (clip_grad_norm_ + AdamW after every step)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--nvtx-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="flat reducer without the per-bucket completion events: every all-reduce is issued after the step")
    ap.add_argument("--only", action="store_true",
                    help="one workload, one line: skip the default run's extra legs (no_graph, eager_baseline, swin_s)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_default(args)


def _release():
    import gc

    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def run_default(args):
    """The driver's default command.  Primary line = `--workload` (vit_b16: the configuration the metric is quoted on); the
    default run (no --only, no profiling flag) adds, in the same JSON line:
      no_graph        the drop-in path as the reference trainer drives it (train.py:265-299): every kernel issued from
                      Python, stock DDP at N > 1, loss.item() per step — no CUDA graph, no flat reducer;
      eager_baseline  the kernel to beat: the oracle port in eager PyTorch (cuBLASLt / ATen) under bf16 autocast on the
                      same GPU, same step (rank 0 only, N = 1 semantics);
      swin_s          the second half of the metric (value / e2e / roofline / clocks of the Swin-S step);
      dino_deit_s     BASELINE config 5 (DINO DeiT-S/16 multi-crop step: teacher forward, student forward / backward, fused
                      DINO loss, EMA) captured as one CUDA graph; at N > 1 the centre all-reduce runs right behind the replay;
      speedup_vs_eager_gpu = value / eager_baseline.value per GPU (the honest speed-up; the CPU ratio is not)."""
    import copy

    import torch.distributed as dist

    extras = not (args.only or args.nvtx_step or args.no_graph or args.workload not in ("vit_b16",))
    if extras:
        args.eager_baseline = True
    line = run_one(args)
    rank = int(os.environ.get("RANK", 0))
    if args.nvtx_step:   # profiling run: one step between cudaProfilerStart / Stop, no line
        if dist.is_initialized():
            dist.destroy_process_group()
        return
    if extras:
        _release()
        a2 = copy.copy(args)
        a2.no_graph, a2.reducer, a2.eager_baseline, a2.e2e_u8 = True, "ddp", False, False
        a2.no_cpu_baseline = a2.no_optimizer_leg = True
        a2.steps = min(args.steps, 10)
        ng = run_one(a2, light=True)
        _release()
        a3 = copy.copy(args)
        a3.workload, a3.batch, a3.eager_baseline, a3.e2e_u8 = "swin_s", 0, True, False
        sw = run_one(a3)
        _release()
        a4 = copy.copy(args)
        a4.workload, a4.batch, a4.eager_baseline, a4.e2e_u8, a4.dino_graph = "dino_deit_s", 0, False, False, True
        a4.no_cpu_baseline = a4.no_optimizer_leg = True
        a4.steps = min(args.steps, 10)
        try:
            dn = run_one(a4, light=True)
        except Exception as exc:  # noqa: BLE001  (never lose the primary line over an extra leg)
            dn = {"error": repr(exc)}
        if rank == 0:
            line["no_graph"] = {k: ng[k] for k in ("value", "unit", "ms_per_step", "e2e", "gpu_launches")}
            line["no_graph"]["host_issue_ms_per_step"] = ng["config"]["host_issue_ms_per_step"]
            line["no_graph"]["reducer"] = ng["config"]["reducer"]
            line["no_graph"]["execution"] = "eager launches from Python (autograd.Function per branch), stock DDP at N > 1"
            line["swin_s"] = {k: sw.get(k) for k in ("value", "unit", "ms_per_step", "e2e", "roofline", "clocks", "gpu_launches",
                                                   "cpu_baseline", "eager_baseline", "with_optimizer", "config")}
            line["dino_deit_s"] = dn if "error" in dn else dict(
                {k: dn.get(k) for k in ("value", "unit", "ms_per_step", "e2e", "clocks", "gpu_launches", "config")},
                crops_per_s=dn["value"] * 10, note="value counts SOURCE images (each = 2 x 224^2 + 8 x 96^2 crops)")
    if rank == 0:
        for d in (line, line.get("swin_s")):
            if d and d.get("eager_baseline"):  # per GPU: the eager leg runs on one GPU
                d["speedup_vs_eager_gpu"] = d["value"] / line["n_gpus"] / d["eager_baseline"]["value"]
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


def run_one(args, light=False):
    """One workload through the product path; returns the JSON line as a dict on rank 0 (None elsewhere)."""
    import torch.distributed as dist
    from vtb200 import dist as vd
    import loss as vloss
    from vtb200 import multi, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (the sm_100a kernels are the only implementation)")
    rank, local_rank, world = vd.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local_rank)
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    W = max(3, args.warmup)

    torch.manual_seed(1234 + rank)
    model = build_model(args.workload).to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    is_dino = args.workload == "dino_deit_s"
    if is_dino:
        if world > 1 and not args.dino_graph:
            args.no_graph = True  # the centre all-reduce sits in the middle of the step: issued eagerly with the rest
        teacher = build_model(args.workload, drop_path=0.0).to(dev)
        teacher.load_state_dict(model.state_dict())
        for p in teacher.parameters():
            p.requires_grad_(False)
        center = torch.zeros(1, 65536, device=dev)
        ema_dst = multi.TensorList([p.detach() for p in teacher.parameters()])
        ema_src = multi.TensorList([p.detach() for p in model.parameters()])
    if not args.no_weight_arena:  # one multi-tensor bf16 cast of all weights per forward instead of one launch per Linear
        multi.enable_weight_arena(model)
        if is_dino:
            multi.enable_weight_arena(teacher)
    net, reducer = model, None
    use_graph = not args.no_graph and not args.nvtx_step
    if world > 1:
        if use_graph:
            args.reducer = "flat"  # the captured graph holds fwd+bwd; gradients are all-reduced right after replay
        if args.reducer == "ddp":
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], broadcast_buffers=False,
                                                            gradient_as_bucket_view=True, bucket_cap_mb=64)
        else:
            # .grad = views into the NCCL buckets; per-bucket "complete" events (external: visible outside the captured
            # graph) let reduce() queue each all-reduce behind its bucket while the rest of the backward pass still runs
            reducer = vd.FlatGradReducer(params, bucket_mb=int(os.environ.get("VTB_BUCKET_MB", 128))).attach(overlap=not args.no_overlap)
    x_dev = torch.randn(B, 3, 224, 224, device=dev)
    y_dev = torch.randint(0, 1000, (B,), device=dev)
    if is_dino:
        crops_dev = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(2)] + \
                    [torch.randn(B, 3, 96, 96, device=dev) for _ in range(8)]
        x_dev = crops_dev

    # --dino-graph at N > 1: the centre update leaves the captured step (see the flag's help)
    dino_state = {"defer_center": bool(is_dino and world > 1 and args.dino_graph and not args.no_graph),
                  "bc": torch.zeros(1, 65536, device=dev) if is_dino else None, "rows": 0}

    def dino_loss_fn(student, teacher_out):
        """DINOLoss.forward (loss.py:119-152): centred / sharpened teacher softmax vs student log-softmax over every
        other crop as ONE fused kernel (vtb_dino_loss: loss + student gradient), then the EMA centre update with its
        all-reduce."""
        from vtb200.blocks import DINOLossFn

        loss = DINOLossFn.apply(student.float(), teacher_out.float(), center, 10, 0.1, 0.04)
        if dino_state["defer_center"]:
            dino_state["bc"].copy_(teacher_out.sum(0, keepdim=True))
            dino_state["rows"] = teacher_out.shape[0]
            return loss
        bc = teacher_out.sum(0, keepdim=True)
        if world > 1:
            dist.all_reduce(bc)
        center.mul_(0.9).add_(bc, alpha=0.1 / (teacher_out.shape[0] * world))  # in place: the buffer is part of the graph
        return loss

    def fwd_bwd(x, y):
        if reducer is not None and reducer.attached:
            reducer.zero()  # one memset per flat bucket; backward accumulates into the bucket views
        else:
            for p in params:
                p.grad = None
        if is_dino:
            with torch.no_grad():
                t_out = teacher(x[:2])
            loss = dino_loss_fn(net(x), t_out)
            loss.backward()
            multi.ema(ema_dst, ema_src, 0.996)  # EMA teacher (train_dino.py:257-261), one launch
            return loss
        loss = vloss.cross_entropy(net(x), y)  # fused log-softmax + NLL + gradient (vtb_mix_loss)
        loss.backward()
        return loss

    graphed = None

    def step(x, y):
        if graphed is not None:
            if x is not x_dev:
                if isinstance(x_dev, (list, tuple)):
                    for dst, src in zip(x_dev, x):
                        dst.copy_(src, non_blocking=True)
                else:
                    x_dev.copy_(x, non_blocking=True)
                    y_dev.copy_(y, non_blocking=True)
            loss = graphed.replay()
        else:
            loss = fwd_bwd(x, y)
        if is_dino and dino_state["defer_center"]:
            dist.all_reduce(dino_state["bc"])
            center.mul_(0.9).add_(dino_state["bc"], alpha=0.1 / (dino_state["rows"] * world))
        if reducer is not None:
            reducer.reduce()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing (`value`)
    if use_graph:
        from vtb200.graph import GraphedStep

        try:
            graphed = GraphedStep(fwd_bwd, (x_dev, y_dev), warmup=2)
        except Exception as exc:  # noqa: BLE001  (a step that cannot be captured is issued eagerly and says so)
            if not is_dino:
                raise
            print(f"bench.py: CUDA-graph capture of the DINO step failed ({exc!r}); issuing eagerly", file=sys.stderr)
            graphed, use_graph = None, False
            dino_state["defer_center"] = False
    graph_grads = [p.grad for p in params] if use_graph else None  # the graph's own (address-stable) gradient buffers
    for _ in range(W):
        step(x_dev, y_dev)
    barrier()
    if args.nvtx_step:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()  # process-wide (backward runs on the autograd thread)
        step(x_dev, y_dev)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps  # host time to ISSUE a step (no sync)
    e1.record()
    barrier()
    ms = vd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    launches = (ops.LAUNCHES - l0)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms * 1e-3)

    # ---------------------------------------------------------------- live roofline pass (dominant kernel: GEMM)
    # one EAGER step with CUDA events around every library call (also counts the launches a replayed graph contains)
    ops.PROFILE = []
    l1 = ops.LAUNCHES
    fwd_bwd(x_dev, y_dev)
    torch.cuda.synchronize()
    launches_per_step = ops.LAUNCHES - l1
    if use_graph:
        launches = launches_per_step * args.steps
    prof, ops.PROFILE = ops.PROFILE, None
    if graph_grads is not None and not (reducer is not None and reducer.attached):
        for p, g in zip(params, graph_grads):  # the eager pass re-created .grad; replays write the captured buffers
            p.grad = g
    gemm = [(n, f, a.elapsed_time(b), nb) for n, f, a, b, nb in prof if n.startswith("gemm_")]
    gemm_ms = sum(t for _, _, t, _ in gemm)
    gemm_flops = sum(f for _, f, _, _ in gemm)
    gemm_bytes = sum(nb for _, _, _, nb in gemm)
    if rank == 0 and not light:
        agg = {}
        for n, f, a, b, _nb in prof:
            t = a.elapsed_time(b)
            d = agg.setdefault(n, [0, 0.0, 0.0])
            d[0] += 1; d[1] += t; d[2] += f
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"breakdown_{args.workload}_n{world}.txt"), "w") as fh:
            tot = sum(v[1] for v in agg.values())
            fh.write(f"# per-op CUDA-event times of ONE instrumented step ({args.workload}); step (uninstrumented) {ms:.2f} ms, sum of ops {tot:.2f} ms\n")
            for n, (c, t, f) in rows:
                tf = f / (t * 1e-3) / 1e12 if f and t > 0 else 0.0
                fh.write(f"{n:44s} n={c:4d} {t:9.3f} ms {100 * t / tot:5.1f}%  {tf:8.1f} TFLOP/s\n")
    peak_tf, peak_hbm, peak_src = peaks()
    # DRAM traffic of the GEMM launches from the committed ncu capture of this same command (per launch, like `achieved`)
    traffic = None
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r[0-9][0-9]_gemm_traffic_{args.workload}.json")))
    tpath = tpaths[-1] if tpaths else None  # the latest round's capture
    if tpath:
        traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    achieved_gbs = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0
    # the GEMM launches of a step are bound by whichever resource they use the larger fraction of: the tensor pipe
    # (ViT-B: K, N >= 768) or HBM (Swin / PVT / Halo: narrow layers, the fp32 residual stream and bf16 activations dominate)
    hbm_bound = achieved_gbs / peak_hbm > achieved / peak_tf
    roofline = {"bound": "hbm" if hbm_bound else "tensor", "kernel": "gemm_tc_kernel (tcgen05)",
                "achieved": achieved_gbs if hbm_bound else achieved, "peak": peak_hbm if hbm_bound else peak_tf,
                "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": (achieved_gbs / peak_hbm) if hbm_bound else (achieved / peak_tf),
                "tensor": {"achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf},
                "hbm": {"achieved": achieved_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": achieved_gbs / peak_hbm,
                        "algorithmic_bytes_per_launch": (gemm_bytes / len(gemm)) if gemm else None},
                "traffic": traffic,
                "traffic_source": "committed ncu capture of this command (not measured in this run)" if tpath else None,
                "achieved_source": "CUDA events around every GEMM launch of ONE instrumented eager step of this run (the "
                                   "graph-timed region replays the same launches)",
                "traffic_unit": "bytes per GEMM launch (ncu dram__bytes_read+write, "
                                + (("profiles/" + os.path.basename(tpath)) if tpath else "no capture committed") + ")",
                "flop_per_launch": (gemm_flops / len(gemm)) if gemm else None, "peak_source": peak_src,
                "launches_per_step": len(gemm), "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / ms,
                "model_tflops": value / world * wl["gflop"] / 1e3, "model_frac": value / world * wl["gflop"] / 1e3 / peak_tf}

    # second half of the BASELINE metric ("attn tensor-pipe %"): the attention launches of the same instrumented step
    # (CUDA events, algorithmic FLOPs / bytes) next to the tensor-pipe activity ncu measured for these kernels
    try:
        att = {k: [(f, a.elapsed_time(b), nb) for n, f, a, b, nb in prof if n == k] for k in ("attention_fwd", "attention_bwd")}
        if att["attention_fwd"]:
            pipe = None
            ppaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_attn_tensor_pipe.json")))
            if ppaths:
                pipe = json.load(open(ppaths[-1])).get(args.workload)
            roofline["attention"] = {"tensor_pipe_pct_ncu": pipe,
                                     "tensor_pipe_source": ("capture profiles/" + os.path.basename(ppaths[-1]) + " (ncu, not measured in "
                                                            "this run; the ms / GB/s / TFLOP/s below ARE measured in this run)") if ppaths else None}
            for k, rows in att.items():
                t = sum(r[1] for r in rows)
                if rows and t > 0:
                    fl, nb = sum(r[0] for r in rows), sum(r[2] for r in rows)
                    roofline["attention"][k] = {"launches": len(rows), "ms_per_step": t, "share_of_step": t / ms,
                                                "tflops": fl / (t * 1e-3) / 1e12, "tensor_frac": fl / (t * 1e-3) / 1e12 / peak_tf,
                                                "gbs": nb / (t * 1e-3) / 1e9, "hbm_frac": nb / (t * 1e-3) / 1e9 / peak_hbm}
    except Exception as exc:  # noqa: BLE001  (reporting only: never lose the bench line over it)
        roofline["attention"] = {"error": repr(exc)}

    # ---------------------------------------------------------------- end-to-end from host buffers (`e2e`)
    e2e = None
    if not args.no_e2e and is_dino:
        crops_host = [c.cpu().pin_memory() for c in crops_dev]

        def e2e_dino(n):
            tot = 0.0
            for _ in range(n):
                xd = [c.to(dev, non_blocking=True) for c in crops_host]
                tot += step(xd, None).item()
            return tot

        e2e_dino(1)
        barrier()
        e0.record()
        e2e_dino(args.steps)
        e1.record()
        barrier()
        ms2 = vd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        e2e = {"value": world * B / (ms2 * 1e-3), "unit": "images/s", "ms_per_step": ms2,
               "h2d_bytes_per_step": sum(c.numel() * 4 for c in crops_host), "d2h_bytes_per_step": 4}
    elif not args.no_e2e:
        n_host = 3
        xs = [torch.randn(B, 3, 224, 224).pin_memory() for _ in range(n_host)]
        ys = [torch.randint(0, 1000, (B,)).pin_memory() for _ in range(n_host)]
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                xd = xs[i % n_host].to(dev, non_blocking=True)
                yd = ys[i % n_host].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return xd, yd, ev

        def e2e_loop(n):
            nxt = prefetch(0)
            tot = 0.0
            for i in range(n):
                xd, yd, ev = nxt
                main_stream.wait_event(ev)
                xd.record_stream(main_stream)
                yd.record_stream(main_stream)
                if i + 1 < n:
                    nxt = prefetch(i + 1)  # overlaps this step's compute, like a DataLoader worker + .to("cuda")
                tot += step(xd, yd).item()  # D2H read of the loss every step (train.py:279)
            return tot

        e2e_loop(2)
        barrier()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        ms2 = vd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        e2e = {"value": world * B / (ms2 * 1e-3), "unit": "images/s", "ms_per_step": ms2,
               "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 8, "d2h_bytes_per_step": 4}

    # ---------------------------------------------------------------- opt-in: e2e through the device input path
    e2e_u8 = None
    if args.e2e_u8 and not is_dino:
        import random as pyrandom

        import device_input as vin

        sampler = vin.MixSampler(0.8, 1.0, 0.25, mix_before_aug=True, rng=pyrandom.Random(rank), noise_seed=rank)
        pipe = vin.DeviceInput(device=dev)  # config/swin-transformer-s.conf:29-31: erasing 0.25, mixup 0.8, cutmix 1.0
        n_host = 3
        srcs = [torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        same = {i: i for i in range(B)}  # partners are drawn inside the batch
        tabs = [vin.pack_table([sampler.sample(i, B, 224, 224) for i in range(B)], same, True, "pixel")
                for _ in range(n_host)]

        def u8_loop(n):
            tot = 0.0
            for i in range(n):
                xd = pipe(srcs[i % n_host], tabs[i % n_host])  # H2D of the uint8 batch + table, one kernel
                tot += step(xd, y_dev).item()
            return tot

        u8_loop(2)
        barrier()
        e0.record()
        u8_loop(args.steps)
        e1.record()
        barrier()
        ms4 = vd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        e2e_u8 = {"value": world * B / (ms4 * 1e-3), "unit": "images/s", "ms_per_step": ms4,
                  "h2d_bytes_per_step": B * 3 * 224 * 224 + B * vin.TABLE_COLS * 4, "d2h_bytes_per_step": 4,
                  "includes": "uint8 H2D + vtb_input_batch (normalise, mixup / cutmix, erasing) + fwd+bwd step, one stream"}

    # ---------------------------------------------------------------- secondary: the same step + optimizer (§8d "also
    # reported included"): clip_grad_norm_ (train.py:294) + AdamW (config/swin-transformer-s.conf:39-42), both multi-tensor
    with_opt = None
    if not args.no_optimizer_leg and not is_dino:
        import optimizer as vopt
        import train_util as vtu

        def wd_skip(name, p):  # factory.py:25-39
            return p.ndim == 1 or any(k in name for k in ("bias", "cls", "norm", "bn", "gain"))

        groups, _ = vtu.add_weight_decay(model.named_parameters(), 0.05, wd_skip)
        opt = vopt.AdamW(list(groups), lr=2.5e-4)

        opt_events = []

        def opt_step():
            loss = step(x_dev, y_dev)
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            vopt.clip_grad_norm_(params, 5.0, defer_to=opt)
            opt.step()
            eb.record()
            opt_events.append((ea, eb))
            return loss

        l2 = ops.LAUNCHES
        opt_step()
        opt_launches = ops.LAUNCHES - l2 - (0 if use_graph else launches_per_step)
        opt_step()
        barrier()
        del opt_events[:]
        e0.record()
        for _ in range(args.steps):
            opt_step()
        e1.record()
        barrier()
        ms3 = vd.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        opt_ms = sum(a.elapsed_time(b) for a, b in opt_events) / len(opt_events)
        with_opt = {"value": world * B / (ms3 * 1e-3), "unit": "images/s", "ms_per_step": ms3,
                    "optimizer_ms_per_step": opt_ms, "optimizer_launches_per_step": opt_launches,
                    "optimizer_bytes_per_step": 32.0 * sum(p.numel() for p in params),
                    "optimizer_gbps": 32.0 * sum(p.numel() for p in params) / (opt_ms * 1e-3) / 1e9,
                    "includes": "fwd+bwd step + clip_grad_norm_(5.0) folded into AdamW(lr 2.5e-4, wd 0.05 / no-decay "
                                "groups): vtb_mt_grad_norm + vtb_mt_adamw"}

    # ---------------------------------------------------------------- opt-in: eager PyTorch on the same GPU (baseline)
    eager = None
    if args.eager_baseline and rank == 0:
        for eb in (B, B // 2, B // 4, B // 8):
            try:
                estep = cpu_reference_step(args.workload, eb, os.cpu_count() or 1, device=dev, autocast=True)
                estep()
                estep()
                torch.cuda.synchronize()
                n = 5
                e0.record()
                for _ in range(n):
                    estep()
                e1.record()
                torch.cuda.synchronize()
                ems = e0.elapsed_time(e1) / n
                eager = {"value": eb / (ems * 1e-3), "unit": "images/s", "ms_per_step": ems, "batch": eb,
                         "kind": "oracle port (plain PyTorch ops: F.linear / matmul / softmax / layer_norm) in eager mode under "
                                 "bf16 autocast, same GPU, same step (fwd + loss + bwd, loss.item() per step)"}
                break
            except torch.OutOfMemoryError:
                estep = None
                torch.cuda.empty_cache()

    # ---------------------------------------------------------------- CPU baseline (oracle port, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = 2 if is_dino else 8
        cstep = cpu_reference_step(args.workload, cb, threads)
        cstep()
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            cstep()
        dt = (time.perf_counter() - t0) / n
        cpu = {"value": cb / dt, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{args.workload} fp32 fwd+bwd, batch {cb}, mean of {n} steps after 1 warm-up (oracle/restate.py)"}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": args.workload, "desc": wl["desc"], "batch_per_gpu": B, "global_batch": B * world,
                           "parallelism": f"dp{world}", "reducer": args.reducer if world > 1 else "none",
                           "allreduce": ("none" if world == 1 else "stock DDP (overlapped)" if reducer is None else
                                         "flat buckets, one all-reduce per bucket after the step" if args.no_overlap else
                                         "flat 128 MiB buckets, all-reduce per bucket behind an external bucket-complete event "
                                         "(overlaps the replaying backward pass)"),
                           "l2": "inputs_exceed_l2 (154 MB batch + multi-GB activations per step >> 126 MB L2)",
                           "timed": "forward + cross-entropy + backward (+ gradient all-reduce); optimizer excluded per metric",
                           "execution": "CUDA graph replay of the captured step" if use_graph else "eager launches",
                           "weight_cast": "per Linear" if args.no_weight_arena else "one multi-tensor launch per forward",
                           "host_issue_ms_per_step": host_ms,
                           "nccl_registered_buckets": bool(reducer is not None and getattr(reducer, "nccl_registered", False))},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "with_optimizer": with_opt}
        if e2e_u8 is not None:
            line["e2e_u8"] = e2e_u8
        if eager is not None:
            line["eager_baseline"] = eager
    if getattr(model, "_vtb_weight_arena", None) is not None:
        multi.disable_weight_arena(model)
    if is_dino and getattr(teacher, "_vtb_weight_arena", None) is not None:
        multi.disable_weight_arena(teacher)
    if world > 1:
        dist.barrier()  # rank 0 ran the baselines alone
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
