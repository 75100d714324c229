"""Drop-in for the reference's models2d.py (class VAE, models2d.py:8-52) running on the B200 kernels.

Same constructor-built submodules (so `state_dict()` keys / shapes equal the reference's: conv1d1..3, fc0, fc11,
fc12, fc2, gru, fc3), same `encode / reparametrize / decode / forward` methods and return values
(`forward(x) -> (probs (B,T,C), mu, logvar)`), usable with the reference's `loss_function` (train.py:31-38),
`loss.backward()` and any torch optimiser.  The nn.Modules only HOLD the parameters; all arithmetic runs in
libmvae_b200.so.  Differences a caller can see: the latent / hidden sizes are constructor arguments (the
reference hard-codes 2 / 501), the input may be u8/long ids (B,T) as well as the reference's float one-hot
(B,T,C), and `precision` selects fp32 check mode or bf16 tensor-core mode.
"""
import threading

import torch
from torch import nn

from .engine import CfgBEngine, param_order

_ENGINE_LOCK = threading.Lock()   # guards the per-device engine caches (DataParallel worker threads)


class _CfgBFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, ids, eps, *params):
        probs, mu, logvar = engine.forward(list(params), ids, eps)
        ctx.engine, ctx.ids, ctx.eps, ctx.generation = engine, ids, eps, engine.generation
        ctx.save_for_backward(*params)
        return probs, mu, logvar

    @staticmethod
    def backward(ctx, dprobs, dmu, dlogvar):
        params = list(ctx.saved_tensors)
        grads = [torch.empty_like(p) for p in params]
        c = lambda t: None if t is None else t.contiguous().float()
        ctx.engine.backward(params, grads, ctx.ids, ctx.eps, c(dprobs), c(dmu), c(dlogvar), generation=ctx.generation)
        return (None, None, None, *grads)


class VAE(nn.Module):
    def __init__(self, latent=2, hidden=501, layers=3, seq_len=120, charset=35, precision="bf16", eps_scale=1.0):
        super().__init__()
        l3 = charset - 26
        self.conv1d1 = nn.Conv1d(seq_len, 9, kernel_size=9)
        self.conv1d2 = nn.Conv1d(9, 9, kernel_size=9)
        self.conv1d3 = nn.Conv1d(9, 10, kernel_size=11)
        self.fc0 = nn.Linear(10 * l3, 435)
        self.fc11 = nn.Linear(435, latent)
        self.fc12 = nn.Linear(435, latent)
        self.fc2 = nn.Linear(latent, latent)
        self.gru = nn.GRU(latent, hidden, layers, batch_first=True)
        self.fc3 = nn.Linear(hidden, charset)
        self.cfg = dict(seq_len=seq_len, charset=charset, latent=latent, hidden=hidden, layers=layers, fc0=435,
                        eps_scale=eps_scale)
        self.precision = precision
        # one live workspace per (module, device).  The dict is shared, not copied, by the replicas
        # nn.DataParallel makes of this module on every forward (train_distributed.py:72: `replica.__dict__` is a shallow
        # copy), so each device's engine is built once and re-used by that device's worker thread.
        self._engines = {}
        self._keys = param_order(layers)
        self.eps_override = None  # tests inject the normal draws here (models2d.py:34 draws them internally)

    # -- plumbing ----------------------------------------------------------------------------------
    def ordered_params(self):
        """Parameters in the C ABI's order, fetched by attribute path: nn.DataParallel's replicas carry their broadcast
        copies as plain tensor attributes (their named_parameters() is empty)."""
        out = []
        for k in self._keys:
            obj = self
            for part in k.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out

    def engine(self, batch, max_len=None):
        dev = self.fc3.weight.device
        key = (batch, self.precision)
        with _ENGINE_LOCK:
            slot = self._engines.get(str(dev))
            if slot is None or slot[0] != key:
                eng = CfgBEngine(batch, precision=self.precision, max_len=float(max_len or self.cfg["seq_len"]),
                                 device=dev, **self.cfg)
                self._engines[str(dev)] = (key, eng)
            else:
                eng = slot[1]
        eng.set_train(self.training)
        return eng

    def _eps(self, batch, device):
        if self.eps_override is not None:
            return self.eps_override.to(device=device, dtype=torch.float32).contiguous()
        return torch.randn(batch, self.cfg["latent"], device=device, dtype=torch.float32)

    # -- reference API (models2d.py:23-52) ----------------------------------------------------------
    def forward(self, x):
        eng = self.engine(x.shape[0])
        ids = eng.to_ids(x)
        eps = self._eps(x.shape[0], ids.device)
        params = [p if p.is_contiguous() else p.contiguous() for p in self.ordered_params()]
        return _CfgBFunction.apply(eng, ids, eps, *params)

    def encode(self, x):
        _, mu, logvar = self.forward(x)
        return mu, logvar

    def reparametrize(self, mu, logvar):
        if self.training:
            return self._eps(mu.shape[0], mu.device) * torch.exp(0.5 * logvar) * self.cfg["eps_scale"] + mu
        return mu

    @torch.no_grad()
    def decode(self, z):
        """probabilities (B,T,C) for latents z (models2d.py:40-47).  INFERENCE ONLY on this path: the result carries no
        autograd graph (training goes through forward() / elbo_step(), which differentiate the whole encode ->
        reparametrize -> decode chain in one call), and it overwrites the activations a pending backward would read."""
        eng = self.engine(z.shape[0])
        params = [p.detach() for p in self.ordered_params()]
        _, probs = eng.decode_greedy(params, z.detach().float().contiguous(), want_probs=True)
        return probs

    @torch.no_grad()
    def decode_greedy(self, z):
        """argmax ids (B,T) u8 of decode(z): train.py:110 / train_sample.py:33 without materialising probs."""
        eng = self.engine(z.shape[0])
        ids, _ = eng.decode_greedy([p.detach() for p in self.ordered_params()], z.detach().float().contiguous())
        return ids

    # -- fused fast path ----------------------------------------------------------------------------
    def elbo_step(self, x, eps=None, max_len=None, use_graph=True):
        """Fused forward + loss_function (train.py:31-38) + backward.  Fills p.grad of every parameter and
        returns the device tensor [loss, max_len*BCE, KL, n_exact_reconstructions].

        With use_graph the ~1.5k kernel launches of a step are captured once into a CUDA graph over static
        input buffers and replayed; the capture is redone if a parameter or gradient tensor moves."""
        eng = self.engine(x.shape[0], max_len)
        if max_len is not None and float(max_len) != eng.desc.max_len:
            eng.desc.max_len = float(max_len)
            eng.destroy_graph()
        ids = eng.to_ids(x)
        if eps is None:
            eps = self._eps(x.shape[0], ids.device)
        else:
            eps = eps.to(ids.device, torch.float32).contiguous()
        params = self.ordered_params()
        for p in params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
        if not use_graph:
            return eng.elbo_step([p.data for p in params], [p.grad for p in params], ids, eps)
        key = (tuple(p.data_ptr() for p in params), tuple(p.grad.data_ptr() for p in params), eng.desc.train)
        if eng._graph is None or getattr(eng, "_graph_key", None) != key:
            eng._ids_static = torch.empty_like(ids)
            eng._eps_static = torch.empty_like(eps)
            eng.capture_elbo_step([p.data for p in params], [p.grad for p in params], eng._ids_static,
                                  eng._eps_static)
            eng._graph_key = key
        eng._ids_static.copy_(ids, non_blocking=True)
        eng._eps_static.copy_(eps, non_blocking=True)
        return eng.launch_graph()


max_len = 120  # script-level global read by loss_function, as in train.py:43 / train_distributed.py:61


def loss_function(recon_x, x, mu, logvar):
    """train.py:31-38 / train_distributed.py:23-30, verbatim semantics (BCE on probabilities times the global
    `max_len`, plus the KL with mu and logvar swapped exactly as shipped)."""
    recon_x = recon_x.contiguous().view(-1)
    x = x.contiguous().view(-1)
    xent_loss = max_len * nn.functional.binary_cross_entropy(recon_x, x, reduction="mean")
    kl_loss = -0.5 * torch.mean(1. + mu - logvar ** 2. - torch.exp(mu))
    return xent_loss + kl_loss
