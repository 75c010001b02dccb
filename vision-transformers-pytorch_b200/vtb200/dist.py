"""Data-parallel plumbing (reference: train.py:102-108 DDP wrap, factory.py:264 per-GPU batch = global // world).

The path shards by images only: every rank holds a full replica and the single exchange step is the
gradient all-reduce.  Two reducers are offered:
  * stock torch DDP (bucketed, overlapped with backward) — what the reference uses;
  * FlatGradReducer — grads packed into a few large flat fp32 buckets, one NCCL all-reduce each (NVLS over
    NVSwitch makes the cost ~size/bandwidth, so few big messages beat many small ones), then unpacked.
Works with gloo on CPU (tests/test_dist_cpu.py) and NCCL on the box.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def per_rank_batch(global_batch, world):
    """factory.py:264."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} not divisible by world size {world}")
    return global_batch // world


def max_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def sum_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


class FlatGradReducer:
    """Mean all-reduce of .grad over ranks through flat fp32 buckets of ~bucket_mb MiB."""

    def __init__(self, params, bucket_mb=128):
        self._events = None       # overlap mode: one "bucket complete" event per bucket (see attach)
        self._hooks = []
        self.comm_stream = None
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, size, cap = [], 0, bucket_mb * (1 << 20) // 4
        for p in self.params:
            if cur and size + p.numel() > cap:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)
        self.attached = False
        self.nccl_registered = False   # buckets live in NCCL-allocated, communicator-registered memory (see _alloc)

    def _alloc(self, sizes, device):
        """One zeroed fp32 buffer per bucket.  On an NCCL job the buffers come from NCCL's own allocator (ncclMemAlloc
        behind torch.cuda.MemPool) and are registered with the communicator: user-buffer registration lets the all-reduce
        run zero-copy over NVLS (the reduction happens in the NVSwitch, the SMs only issue multimem loads / stores) instead
        of staging every chunk through NCCL's internal buffers.  Opt-in (VTB_NCCL_POOL=1): measured on 2 and 8 B200s it
        changes nothing for this step (ViT-B, 343 MB of gradients: 35.37 - 35.50 ms per step at N=8 either way,
        profiles/r02_ab_n8_nccl_registered_buckets.log) — the collective is already hidden behind the backward pass and
        what it costs is the power / HBM share it takes from the GEMMs, not its own duration.  Otherwise, or on a non-NCCL
        backend, one rank or any failure of the allocator / registration: plain torch.zeros."""
        use = (device.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
               and dist.get_world_size() > 1 and os.environ.get("VTB_NCCL_POOL", "0") == "1")
        if use:
            try:
                backend = dist.group.WORLD._get_backend(torch.device("cuda"))
                dist.barrier()  # the communicator must exist before memory can be registered with it
                pool = torch.cuda.MemPool(backend.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    flats = [torch.zeros(n, dtype=torch.float32, device=device) for n in sizes]
                backend.register_mem_pool(pool)
                self._pool, self.nccl_registered = pool, True   # keep the pool alive as long as the buckets
                return flats
            except Exception as e:  # noqa: BLE001  (allocator not built in, no cuMem / multicast support, ...)
                import warnings

                warnings.warn(f"vtb200.dist: NCCL-registered gradient buckets unavailable ({e!r}); using plain device memory")
        return [torch.zeros(n, dtype=torch.float32, device=device) for n in sizes]

    def attach(self, overlap=False):
        """Make every parameter's .grad a VIEW into its flat bucket: backward then accumulates straight into the
        buffers NCCL reduces (no pack / unpack passes, no per-parameter kernels), `zero()` is one memset per bucket
        and `reduce()` one all-reduce + one scale per bucket.  Call `zero()` instead of setting .grad = None.

        overlap=True (what stock DDP does, train.py:103-107, but compatible with a captured step): a post-accumulate hook
        per parameter counts the bucket's gradients and records a "bucket complete" EVENT behind the last one.  The events
        are `external` ones: captured into a CUDA graph they become event-record nodes that streams OUTSIDE the graph can
        wait on, so `reduce()`, called right after `graph.replay()` was launched, queues each bucket's all-reduce on a side
        stream behind its event and the collectives run while the rest of the backward pass is still replaying.  No NCCL call
        is captured.  Every parameter must receive a gradient in every step (DDP's find_unused_parameters=False contract)."""
        flats = self._alloc([sum(p.numel() for p in bucket) for bucket in self.buckets], self.buckets[0][0].device)
        for i, bucket in enumerate(self.buckets):
            flat = flats[i]
            off = 0
            for p in bucket:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self._flat[i] = flat
        self.attached = True
        if overlap:
            self._arm_overlap()
        return self

    def _new_event(self):
        return torch.cuda.Event(external=True)

    def _arm_overlap(self):
        cuda = self._flat[0].is_cuda
        self._events = [self._new_event() if cuda else None for _ in self.buckets]
        self._seen = [0] * len(self.buckets)
        self._order = []          # bucket indices in the order they completed during the last backward
        self.comm_stream = torch.cuda.Stream(device=self._flat[0].device) if cuda else None

        def make(i, n):
            def hook(_p):
                self._seen[i] += 1
                if self._seen[i] == n:
                    self._seen[i] = 0
                    if i in self._order:
                        self._order.remove(i)
                    self._order.append(i)
                    if self._events[i] is not None:
                        self._events[i].record()   # on the current (possibly capturing) stream
            return hook

        for i, bucket in enumerate(self.buckets):
            for p in bucket:
                self._hooks.append(p.register_post_accumulate_grad_hook(make(i, len(bucket))))

    def detach_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks, self._events = [], None

    def _views(self):
        for flat, bucket in zip(self._flat, self.buckets):
            off = 0
            for p in bucket:
                yield p, flat[off:off + p.numel()].view_as(p)
                off += p.numel()

    def reattach(self):
        """Repair the .grad -> bucket link.  `p.grad = None` (train_util.cancel_last_layer_grad on the DINO `last` layer,
        optimizer.zero_grad(set_to_none=True)) or a re-assigned .grad makes autograd accumulate OUTSIDE the bucket; the
        all-reduce would then average a stale slice and the replicas would silently diverge.  Called by zero() and
        reduce(): a detached gradient is copied into its slice (None = no gradient this step = zeros) and re-pointed.
        Returns the number of parameters that had to be repaired."""
        n = 0
        for p, view in self._views():
            g = p.grad
            if g is not None and g.data_ptr() == view.data_ptr() and g.shape == view.shape:
                continue
            n += 1
            with torch.no_grad():
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
            p.grad = view
        return n

    def zero(self):
        for flat in self._flat:
            flat.zero_()
        if self.attached:
            self.reattach()

    def reduce(self):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        if self.attached and self._events is not None:
            # overlap mode: one all-reduce per bucket behind its "bucket complete" event, in completion order
            order = self._order if len(self._order) == len(self.buckets) else list(reversed(range(len(self.buckets))))
            if self.comm_stream is None:   # CPU / gloo: same logic, no streams
                for i in order:
                    dist.all_reduce(self._flat[i])
                    self._flat[i].mul_(1.0 / world)
                return
            main = torch.cuda.current_stream(self._flat[0].device)
            avg = dist.get_backend() == "nccl"   # NCCL averages inside the collective: no extra pass over the bucket
            for i in order:
                self.comm_stream.wait_event(self._events[i])
                with torch.cuda.stream(self.comm_stream):
                    if avg:
                        dist.all_reduce(self._flat[i], op=dist.ReduceOp.AVG, async_op=True).wait()
                    else:
                        dist.all_reduce(self._flat[i], async_op=True).wait()
                        self._flat[i].mul_(1.0 / world)
            main.wait_stream(self.comm_stream)   # the next step's zero() / the optimizer see reduced gradients
            return
        if self.attached:
            self.reattach()
            works = [dist.all_reduce(flat, async_op=True) for flat in self._flat]
            for work, flat in zip(works, self._flat):
                work.wait()
                flat.mul_(1.0 / world)
            return
        works = []
        for i, bucket in enumerate(self.buckets):
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            n = sum(g.numel() for g in grads)
            if self._flat[i] is None or self._flat[i].numel() != n or self._flat[i].device != grads[0].device:
                self._flat[i] = torch.empty(n, dtype=torch.float32, device=grads[0].device)
            flat = self._flat[i]
            torch.cat([g.reshape(-1) for g in grads], out=flat)
            works.append((dist.all_reduce(flat, async_op=True), flat, bucket))
        for work, flat, bucket in works:
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p) / world
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
