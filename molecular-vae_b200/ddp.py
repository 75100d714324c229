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
    on NCCL's stream (it waits for the work enqueued so far), launch phase p+1 meanwhile; finish() joins.  The mean over
    the ranks is taken by the collective itself (NCCL's AVG reduction) -- no separate pass over the flat buffer; backends
    without AVG (gloo, the CPU tests) sum and divide."""

    def __init__(self, flat, buckets, group=None):
        self.flat, self.buckets, self.group = flat, buckets, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._avg = self.world > 1 and dist.get_backend(group) == "nccl"
        self._pending = []

    def after_phase(self, p):
        if self.world == 1:
            return
        lo, hi = self.buckets[p]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._pending.append(dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True))

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        if self.world > 1 and not self._avg:
            self.flat.div_(self.world)


class GraphedDataParallelStep:
    """The whole data-parallel step as ONE CUDA graph: the phases of the fused step and, forked off after each phase, the
    all-reduce (NCCL AVG, captured on NCCL's stream) of the gradient bucket that phase finalised; the branches join at the end.
    One graph launch per step -- no host work between the phases, no separate division pass.  `run_phase(p)` must enqueue
    phase p on the current stream over FIXED buffers (e.g. CfgBEngine.elbo_step_phase)."""

    def __init__(self, run_phase, n_phases, flat, buckets, group=None, warmup=True):
        self.reducer = PhasedAllReduce(flat, buckets, group)
        self.n_phases = n_phases
        if warmup:                       # outside capture: lazy allocations, opt-in shared memory, NCCL communicator set-up
            for p in range(n_phases):
                run_phase(p)
                self.reducer.after_phase(p)
            self.reducer.finish()
        torch.cuda.synchronize()
        from ._lib import lib
        n0 = lib.mvae_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            for p in range(n_phases):
                run_phase(p)
                self.reducer.after_phase(p)
            self.reducer.finish()
        self.launches_per_step = int(lib.mvae_launch_count() - n0)   # this library's kernels / memsets inside the graph

    def step(self):
        self.graph.replay()


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


def moses_readiness_order(keys, d_layers, head_keys=()):
    """Backward-readiness order of the MOSES VAE's parameters (`keys` in C-ABI order, mosesvae.moses_param_order /
    mosesfile.mosesfile_param_order; `head_keys` = the property head's parameters) and, per phase of
    mvae_moses_step_ex / mvae_moses_joint_step, how many leading keys of that order are final after the phase:
      phase 0   decoder_fc, the top decoder layer, the property head
      phase k   decoder layer d_layers-1-k
      last      decoder layer 0, decoder_lat, the mu / logvar heads, the encoder GRU(s), x_emb
    A FlatGradBuffer laid out in this order makes every phase's bucket one contiguous slice."""
    layer = lambda l: [k for k in keys if k.startswith("decoder_rnn.") and k.endswith(f"_l{l}")]
    order, cuts = [], []
    top = d_layers - 1
    order += [k for k in keys if k.startswith("decoder_fc.")] + layer(top) + list(head_keys)
    for p in range(d_layers):
        if p > 0:
            order += layer(top - p)
        if p == d_layers - 1:
            order += [k for k in keys if k not in order]
        cuts.append(len(order))
    return order, cuts


class MosesPhasedStep:
    """Data-parallel step of the MOSES VAE (optionally with the property head) over FIXED buffers: one CUDA graph per phase
    of the fused step, the gradient bucket a phase finalises is all-reduced on NCCL's stream while the next phase's BPTT sweep
    runs (the reference's only multi-GPU mechanism for this model is the commented-out DDP of moses_train_distrib.py:30-42,191).
    Loss semantics: KL and the property loss are means over the GLOBAL batch, the reconstruction CE a mean over
    the GLOBAL count of non-pad targets, so rank r runs with kl_weight * N B_r / B, recon_weight * N M_r / M and
    binding_weight * N B_r / B (moses_rank_weights) and the exchange averages."""

    def __init__(self, model, x, eps, kl_weight=1.0, recon_weight=1.0, binding=None, binding_weight=1.0, dropout=None,
                 group=None, single_graph=True):
        self.model, self.group = model, group
        _, self.ids, self.lens = model._pack(x)
        dev = self.ids.device
        self.eps = eps.to(dev, torch.float32).contiguous()
        B = self.ids.shape[0]
        M = int(self.lens._host_copy.sum()) - B
        ks, rs = moses_rank_weights(B, M, group)
        self.kl_weight, self.recon_weight = kl_weight * ks, recon_weight * rs
        self.dropout = dropout if dropout is not None else model._dropout()
        L = model.cfg["d_layers"]
        named = dict(model.named_parameters())
        head = getattr(model, "binding_model", None) if binding is not None else None
        head_keys = [f"binding_model.binding_model.{k}" for k in head.KEYS] if head is not None else []
        order, cuts = moses_readiness_order(model._keys, L, head_keys)
        self.gbuf = FlatGradBuffer([named[k] for k in order])
        numel = [named[k].numel() for k in order]
        ends = [sum(numel[:c]) for c in cuts]
        self.buckets = [(0 if i == 0 else ends[i - 1], ends[i]) for i in range(L)]
        self.params = [named[k].data for k in model._keys]
        self.grads = [named[k].grad for k in model._keys]
        self.joint = None
        if head is not None:
            self.target = binding.to(dev, torch.float32).contiguous().view(-1)
            self.joint = (head, self.target, binding_weight * ks, [named[k].grad for k in head_keys])
        self.reducer = PhasedAllReduce(self.gbuf.flat, self.buckets, group)
        self.graphs = []
        self.single = None
        self._run_phase(-1)                      # warm-up outside capture: workspaces, scratch, opt-in shared memory
        torch.cuda.synchronize()
        if single_graph:
            # one graph for the whole step: phases + the bucket all-reduces as forked branches (GraphedDataParallelStep)
            self.single = GraphedDataParallelStep(self._run_phase, L, self.gbuf.flat, self.buckets, group, warmup=self.reducer.world > 1)
            return
        for p in range(L):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_phase(p)
            self.graphs.append(g)

    def _run_phase(self, p):
        self.model._run(self.params, self.grads, self.ids, self.lens, self.eps, self.kl_weight, self.recon_weight, False,
                        dropout=self.dropout, phase=p, joint=self.joint)

    def step(self):
        if self.single is not None:
            self.single.step()
            return self.model._last_scalars
        for p, g in enumerate(self.graphs):
            g.replay()
            self.reducer.after_phase(p)
        self.reducer.finish()
        return self.model._last_scalars
