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


def phase_buckets(keys, numels, layers):
    """Contiguous [lo, hi) element ranges of the flat gradient buffer that become final after each phase of
    mvae_cfgb_elbo_step_phase (`keys` / `numels` in state_dict = C-ABI order): phase 0 -> top GRU layer + fc3,
    phase k -> GRU layer L-1-k, phase L-1 -> additionally everything in front of the GRU (encoder + latent layers).
    The ranges tile the buffer exactly once."""
    offs, off = {}, 0
    for k, n in zip(keys, numels):
        offs[k] = off
        off += n
    total = off
    starts = [offs[f"gru.weight_ih_l{l}"] for l in range(layers)]
    out = []
    for p in range(layers):
        l = layers - 1 - p
        lo = 0 if l == 0 else starts[l]
        hi = total if p == 0 else starts[l + 1]
        out.append((lo, hi))
    return out


class PhasedAllReduce:
    """Bucketed gradient exchange overlapped with the backward sweep: launch phase p, start the all-reduce of its bucket
    on NCCL's stream (it waits for the work enqueued so far), launch phase p+1 meanwhile; finish() joins and averages."""

    def __init__(self, flat, buckets, group=None):
        self.flat, self.buckets, self.group = flat, buckets, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._pending = []

    def after_phase(self, p):
        if self.world == 1:
            return
        lo, hi = self.buckets[p]
        self._pending.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self.world > 1:
            self.flat.div_(self.world)


def shard_rows(n_rows, rank, world):
    """Rows [lo, hi) of a global batch owned by `rank` (equal shards, remainder to the low ranks)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sorted(seqs, rank, world):
    """Shard of a length-sorted list of sequences (the pack_sequence contract of mosesvae.py:151): every world-th sequence,
    so each shard stays sorted and the shards' length distributions (= work) match."""
    return list(seqs[rank::world])


def moses_rank_weights(n_sequences, n_targets, group=None):
    """Per-rank scales (kl_scale, recon_scale) for the MOSES VAE step under data parallelism: the reference
    computes KL as a mean over the GLOBAL batch (mosesvae.py:162) and the reconstruction CE as a mean over the GLOBAL count of
    non-pad targets (mosesvae.py:193-197).  With rank r holding B_r sequences and M_r targets,
        kl_global = sum_r (B_r / B) kl_r,   recon_global = sum_r (M_r / M) recon_r,
    so a step run with kl_weight * kl_scale and recon_weight * recon_scale, followed by the usual gradient all-reduce MEAN over
    the N ranks, yields exactly the full-batch gradient: kl_scale = N B_r / B, recon_scale = N M_r / M (both 1 for equal
    shards with equal target counts).  Two scalars are all-reduced; the tensors stay on the host."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0, 1.0
    world = dist.get_world_size(group)
    t = torch.tensor([float(n_sequences), float(n_targets)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tot_b, tot_m = (float(v) for v in t.cpu())
    return world * float(n_sequences) / tot_b, world * float(n_targets) / tot_m
