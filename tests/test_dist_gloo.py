"""CPU, world_size 2, gloo: the host-side multi-GPU logic (flat-gradient all-reduce, frame sharding, occupancy broadcast)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "lzzx-nerf_b200"))
    from b2nerf.dist import FlatGradBuffer, shard_frames, gather_order
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                 # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(7, 5, bias=False), torch.nn.ReLU(), torch.nn.Linear(5, 3, bias=False))
    table = torch.nn.Parameter(torch.randn(40, 1))
    params = [table] + list(net.parameters())
    fg = FlatGradBuffer(params)
    fg.check_attached()
    torch.manual_seed(100 + rank)                        # different data per rank (data-parallel over rays)
    x = torch.randn(16, 7); idx = torch.randint(0, 40, (16,))
    fg.zero_()
    loss = (net(x) * table[idx]).pow(2).mean()
    loss.backward()
    fg.check_attached()                                  # autograd accumulated INTO the flat views
    local = fg.flat.clone()
    fg.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok_mean = torch.allclose(fg.flat, torch.stack(gathered).mean(0), atol=1e-7)
    ok_views = torch.equal(table.grad.view(-1), fg.flat[:40])
    # identical optimizer step on every rank afterwards
    opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.0, 0.99))
    opt.step()
    w = torch.cat([p.detach().view(-1) for p in params])
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    ok_sync = all(torch.equal(ws[0], t) for t in ws)
    # inf on one rank reaches every rank through the sum (GradScaler skips consistently)
    fg.zero_()
    if rank == 1:
        fg.flat[3] = float("inf")
    fg.all_reduce_mean()
    ok_inf = bool(torch.isinf(fg.flat[3]))
    # frame sharding
    shards = [shard_frames(11, r, world) for r in range(world)]
    ok_shard = sorted(sum(shards, [])) == list(range(11)) and shards[rank] == list(range(rank, 11, world))
    inv = gather_order(11, world)
    cat = sum(shards, [])
    ok_order = [cat[inv[f]] for f in range(11)] == list(range(11))
    # occupancy broadcast
    bf = torch.full((64,), rank + 1, dtype=torch.uint8)
    dist.broadcast(bf, src=0)
    ok_bcast = bool((bf == 1).all())
    q.put((rank, ok_mean, ok_views, ok_sync, ok_inf, ok_shard, ok_order, ok_bcast))
    dist.destroy_process_group()


def test_flat_grad_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r
