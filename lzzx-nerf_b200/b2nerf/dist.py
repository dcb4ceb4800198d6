"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference has no working multi-GPU path (a vestigial DDP wrap, TrainerUtil.py:79-81), so this is new:
  * inference: frames are independent -> round-robin frame sharding, no collective;
  * training: data-parallel over rays with ONE all-reduce per step over a single flat fp32 gradient buffer (tables + MLPs + audio nets +
    individual codes = 683 509 floats = 2.73 MB): every parameter's .grad is a view into that buffer, so autograd accumulates in place
    and no packing copy is needed;
  * the occupancy bitfield is refreshed on rank 0 (update_extra_state uses unseeded RNG, renderer.py:707,747) and broadcast (256 KB).
"""
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


class FlatGradBuffer:
    """One contiguous fp32 buffer aliasing the .grad of every trainable parameter; all-reduced once per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
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
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
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
