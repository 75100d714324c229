"""Data-parallel plumbing for one-process-per-GPU training (replaces the reference's single-process
nn.DataParallel, train_distributed.py:72): every parameter's .grad is a view into ONE flat fp32 buffer, so a step's
gradient exchange is a single all-reduce; dividing by the world size reproduces DataParallel's full-batch mean loss
for equal shards (train_distributed.py:87-89)."""
import torch
import torch.distributed as dist


class FlatGradBuffer:
    def __init__(self, params):
        self.params = list(params)
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def grads(self):
        return [p.grad for p in self.params]

    def allreduce_mean(self, group=None, async_op=False):
        """Sum over ranks then divide by the world size (no-op without an initialised process group)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work, world
        self.flat.div_(world)
        return None


def shard_rows(n_rows, rank, world):
    """Rows [lo, hi) of a global batch owned by `rank` (equal shards, remainder to the low ranks)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
