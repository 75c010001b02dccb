"""world_size-2 gloo test of the data-parallel plumbing (runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from vtb200 import dist as vd

    r, lr, w = vd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    red = vd.FlatGradReducer(model.parameters(), bucket_mb=1)
    x = torch.full((2, 8), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    red.reduce()
    gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
    want = sum(gathered) / world
    got = torch.cat([p.grad.flatten() for p in model.parameters()])
    ok = torch.allclose(got, want, atol=1e-6)
    # attached mode: .grad are views into the flat buckets, backward accumulates into them, reduce() is in place
    model2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    model2.load_state_dict(model.state_dict())
    red2 = vd.FlatGradReducer(model2.parameters(), bucket_mb=1).attach()
    for _ in range(2):  # second round checks zero() really clears the views
        red2.zero()
        model2(x).sum().backward()
        red2.reduce()
    got2 = torch.cat([p.grad.flatten() for p in model2.parameters()])
    ok = ok and torch.allclose(got2, want, atol=1e-6)
    ok = ok and all(p.grad.data_ptr() >= red2._flat[0].data_ptr() for p in red2.buckets[0])
    # overlap mode (per-bucket completion hooks; on the CPU the events / side stream are absent, the bookkeeping is the same)
    model4 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    model4.load_state_dict(model.state_dict())
    red4 = vd.FlatGradReducer(model4.parameters(), bucket_mb=1)
    red4.buckets = [[p] for p in red4.params]          # one bucket per parameter: exercises the completion order
    red4._flat = [None] * len(red4.buckets)
    red4.attach(overlap=True)
    for _ in range(2):
        red4.zero()
        model4(x).sum().backward()
        order = list(red4._order)
        red4.reduce()
    got4 = torch.cat([p.grad.flatten() for p in model4.parameters()])
    ok = ok and torch.allclose(got4, want, atol=1e-6) and sorted(order) == list(range(6)) and order[0] >= 4
    # a detached .grad (train_util.cancel_last_layer_grad sets p.grad = None on the DINO `last` layer; so does
    # optimizer.zero_grad(set_to_none=True)) must not leave the replicas averaging a stale bucket slice
    import train_util as T

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(8, 8)
            self.last = torch.nn.Linear(8, 4)

        def forward(self, t):
            return self.last(self.body(t))

    torch.manual_seed(1)
    net = Net()
    red3 = vd.FlatGradReducer(net.parameters(), bucket_mb=1).attach()
    red3.zero()
    net(x).sum().backward()
    red3.reduce()
    T.cancel_last_layer_grad(0, net, freeze=1)              # after the all-reduce, as train_dino.py orders it
    ok = ok and net.last.weight.grad is None and net.body.weight.grad is not None
    red3.zero()                                             # next step: the link is repaired, not silently lost
    ok = ok and all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in red3._views())
    net(x).sum().backward()
    local3 = torch.cat([p.grad.flatten().clone() for p in net.parameters()])
    for p in net.last.parameters():                         # detached BEFORE reduce(): copied back into the bucket
        p.grad = p.grad.clone()
    repaired = red3.reattach()
    red3.reduce()
    g3 = [torch.zeros_like(local3) for _ in range(world)]
    dist.all_gather(g3, local3)
    got3 = torch.cat([p.grad.flatten() for p in net.parameters()])
    ok = ok and repaired == 2 and torch.allclose(got3, sum(g3) / world, atol=1e-6)
    mx = vd.max_over_ranks(rank + 10, "cpu")
    sm = vd.sum_over_ranks(rank + 1, "cpu")
    q.put((rank, bool(ok), mx, sm, vd.per_rank_batch(256, world)))
    dist.destroy_process_group()


def test_flat_reducer_and_rank_helpers_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True, 11.0, 3.0, 128), (1, True, 11.0, 3.0, 128)]


def test_bucketing_covers_every_parameter_once():
    from vtb200 import dist as vd

    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 300000, 5, 700000, 1)]
    red = vd.FlatGradReducer(ps, bucket_mb=1)
    flat = [id(p) for b in red.buckets for p in b]
    assert flat == [id(p) for p in ps] and len(red.buckets) >= 3
