"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference has no working multi-GPU path (a vestigial DDP wrap, TrainerUtil.py:79-81), so this is new:
  * inference: frames are independent -> round-robin frame sharding, no collective;
  * training: data-parallel over rays with ONE all-reduce per step over a single flat fp32 gradient buffer (tables + MLPs + audio nets +
    individual codes = 683 509 floats = 2.73 MB): every parameter's .grad is a view into that buffer, so autograd accumulates in place
    and no packing copy is needed;
  * the occupancy bitfield is refreshed on rank 0 (update_extra_state uses unseeded RNG, renderer.py:707,747) and broadcast (256 KB).
"""
import ctypes
import os

import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """Frame f is rendered by rank f mod world; returned in increasing order."""
    return list(range(rank, n_frames, world))


def gather_order(n_frames, world):
    """Position of every frame in the concatenation of the per-rank shards -> permutation that restores frame order."""
    order = [f for r in range(world) for f in shard_frames(n_frames, r, world)]
    inv = [0] * n_frames
    for pos, f in enumerate(order):
        inv[f] = pos
    return inv


class _DevMem:
    """A raw device allocation presented through __cuda_array_interface__ (torch.as_tensor aliases it and keeps this object alive)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = dict(shape=(int(n_floats),), typestr="<f4", data=(int(ptr), False), version=2)


class PeerComm:
    """The flat gradient buffer as NVLink peer memory + the one-kernel all-reduce over it (csrc/peer_allreduce.cu): every rank's buffer is a cudaMalloc block whose
    CUDA IPC handle is exchanged through torch.distributed and opened on all peers (one process per GPU, one node)."""

    def __init__(self, n_floats, device, group=None):
        from ._lib import lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = int(n_floats)
        self.bytes = (self.n * 4 + 15) // 16 * 16
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        lib().call("b2n_peer_alloc", self.bytes, ctypes.byref(ptr), handle)
        self.ptr = ptr
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.comm = ctypes.c_void_p()
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * self.world)
        lib().call("b2n_peer_comm_create", ctypes.byref(self.comm), self.rank, self.world, ptr, blob, self.bytes)
        self.flat = torch.as_tensor(_DevMem(ptr.value, self.bytes // 4), device=device)[:self.n]
        dist.barrier(group=group)                      # every rank has opened every buffer before anyone launches

    def all_reduce_mean(self):
        from ._lib import lib
        lib().call("b2n_peer_allreduce_mean", self.comm, self.n, torch.cuda.current_stream(self.flat.device).cuda_stream)

    def error(self):
        """Non-zero if a barrier timed out (some rank did not make the matching call).  Synchronises."""
        from ._lib import lib
        v = ctypes.c_int32()
        lib().call("b2n_peer_error", self.comm, ctypes.byref(v), torch.cuda.current_stream(self.flat.device).cuda_stream)
        return int(v.value)


def _try_peer_comm(n_floats, device, group=None):
    """PeerComm when every rank can set it up (CUDA, > 1 rank, one node with peer access, not disabled by B2N_PEER_ALLREDUCE=0); else None on EVERY rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and device.type == "cuda"):
        return None
    ok, comm = os.environ.get("B2N_PEER_ALLREDUCE", "1") != "0", None
    if ok:
        try:
            comm = PeerComm(n_floats, device, group)
        except Exception as e:                            # noqa: BLE001 — IPC / peer access unavailable: every rank falls back together
            ok, comm = False, None
            PeerComm.last_error = repr(e)
    flag = torch.tensor([1.0 if ok else 0.0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return comm if float(flag.item()) == 1.0 else None


class FlatGradBuffer:
    """One contiguous fp32 buffer aliasing the .grad of every trainable parameter; all-reduced once per step — by the peer-memory kernel of
    csrc/peer_allreduce.cu when the job runs one process per GPU on one NVLink node (`self.peer`), else by torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, params, peer=True):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.peer = _try_peer_comm(total, ref.device) if peer else None
        self.flat = self.peer.flat.zero_() if self.peer is not None else torch.zeros(total, dtype=torch.float32, device=ref.device)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError("FlatGradBuffer: parameters must be float32")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()          # one memset instead of one per parameter; the views stay attached

    def check_attached(self):
        """optimizer.zero_grad(set_to_none=True) would detach the views — call zero_() instead; this asserts the invariant."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * 4:
                raise RuntimeError("FlatGradBuffer: a parameter's .grad no longer aliases the flat buffer")
            off += p.numel()

    def all_reduce_mean(self, group=None):
        """Sum over ranks then divide by world size (== DistributedDataParallel's gradient averaging).  An inf/nan produced by any rank
        propagates through the sum, so every rank's GradScaler skips the same steps without a second collective."""
        if self.peer is not None:
            self.peer.all_reduce_mean()                # one kernel: barrier, two-shot reduction over NVLink peer memory (mean), barrier
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat

    def nbytes(self):
        return self.flat.numel() * 4


def broadcast_occupancy(model, src=0, group=None):
    """Keep density grid / bitfield identical on all ranks after rank `src` refreshed them."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model.density_bitfield, src=src, group=group)
        dist.broadcast(model.density_grid, src=src, group=group)
