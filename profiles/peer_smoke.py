#!/usr/bin/env python
"""2-GPU smoke of the IPC peer all-reduce with progress prints (run under torchrun, wrapped in `timeout`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
def say(*a):
    print(f"[rank {rank} {time.time() % 1000:.2f}]", *a, flush=True)
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
say("pg up")
t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); say("nccl ok", t[0].item())
from b2nerf.dist import PeerComm
try:
    pc = PeerComm(683509, dev)
    say("PeerComm ok")
except Exception as e:
    say("PeerComm FAILED", repr(e)); raise
x = torch.full((683509,), float(rank + 1), device=dev)
pc.flat.copy_(x); torch.cuda.synchronize(); dist.barrier(); say("filled")
pc.all_reduce_mean(); torch.cuda.synchronize(); say("allreduce done", pc.flat[:3].tolist(), pc.flat[-3:].tolist(), "err", pc.error())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); e0.record()
for _ in range(100):
    pc.all_reduce_mean()
e1.record(); torch.cuda.synchronize(); say("100 calls:", e0.elapsed_time(e1) / 100 * 1e3, "us each, err", pc.error())
y = torch.randn(683509, device=dev)
dist.all_reduce(y); torch.cuda.synchronize(); dist.barrier(); e0.record()
for _ in range(100):
    dist.all_reduce(y)
e1.record(); torch.cuda.synchronize(); say("nccl 100 calls:", e0.elapsed_time(e1) / 100 * 1e3, "us each")
dist.destroy_process_group()
