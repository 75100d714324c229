"""Drop-in for the reference's models.py (models.py:6-165) running on the B200 kernels.

Same classes and constructor signatures (`MolecularVAE(i=120, o=292, c=35)`, `MolEncoder`, `MolDecoder`, `Lambda`, the
helper modules `Flatten`, `Repeat`, `TimeDistributed`, `SELU`, `ConvSELU`), same submodule tree, hence identical
`state_dict()` keys and shapes.  As in the reference the recurrent attributes are called `gru` but are
`nn.LSTM` (models.py:117,156).  `MolecularVAE.forward(LongTensor[B,i]) -> (probs[B,i,c], mu, logvar)` (models.py:104-106),
`model.decoder(z)` (train_sample.py:32) and the `encoder.lmbd.mu / .log_v` side effects (models.py:90-91) are kept.
The nn.Modules only HOLD the parameters; all arithmetic runs in libmvae_b200.so (cfga.cu); there is no CPU fallback.
`elbo_step()` is the fused fast path for train.py:98-101.
"""
import ctypes
import threading

import torch
from torch import nn

from . import _lib
from ._lib import CfgADesc, check, lib
from .engine import _p, _ptr_table, _stream

_ENGINE_LOCK = threading.Lock()   # guards the per-device engine caches (DataParallel worker threads)


def cfga_param_order(enc_layers=3, dec_layers=4):
    """state_dict keys in the order of the C ABI's pointer table (include/mvae_b200.h)."""
    keys = ["encoder.embedding.weight"]
    for l in range(enc_layers):
        keys += [f"encoder.gru.weight_ih_l{l}", f"encoder.gru.weight_hh_l{l}", f"encoder.gru.bias_ih_l{l}",
                 f"encoder.gru.bias_hh_l{l}"]
    for n in ("conv_1.0", "conv_2.0", "conv_3.0", "dense_1.0", "lmbd.z_mean", "lmbd.z_log_var"):
        keys += [f"encoder.{n}.weight", f"encoder.{n}.bias"]
    keys += ["decoder.latent_input.0.weight", "decoder.latent_input.0.bias"]
    for l in range(dec_layers):
        keys += [f"decoder.gru.weight_ih_l{l}", f"decoder.gru.weight_hh_l{l}", f"decoder.gru.bias_ih_l{l}",
                 f"decoder.gru.bias_hh_l{l}"]
    return keys + ["decoder.decoded_mean.module.0.weight", "decoder.decoded_mean.module.0.bias"]


class CfgAEngine:
    """One instance per (shape, precision): owns the HBM workspace and calls the C ABI on torch's current stream."""

    def __init__(self, batch, seq_len=120, charset=35, embed=30, enc_hidden=72, enc_layers=3, latent=292, dec_hidden=1024,
                 dec_layers=4, precision="bf16", max_len=120.0, eps_scale=1e-2, device=None):
        if not torch.cuda.is_available():
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        self.desc = CfgADesc(batch, seq_len, charset, embed, enc_hidden, enc_layers, latent, dec_hidden, dec_layers, prec,
                             float(max_len), float(eps_scale))
        self.ws_bytes = lib.mvae_cfga_workspace_bytes(ctypes.byref(self.desc))
        if self.ws_bytes == 0:
            raise ValueError("invalid / unsupported Config-A description")
        with torch.cuda.device(self.device):
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
        self._ws_ptr = ctypes.c_void_p(self.ws.data_ptr() + (-self.ws.data_ptr()) % 256)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._graph = None
        self._graph_keep = None
        self.generation = 0   # bumped by every call that rewrites the workspace (see CfgBEngine.generation)

    def _args(self):
        return self._ws_ptr, self.ws_bytes, _stream()

    def check_device_error(self):
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            check(lib.mvae_cfga_read_error(ctypes.byref(self.desc), self._ws_ptr, self.ws_bytes, ctypes.byref(flag), _stream()))
        if flag.value:
            raise _lib.MvaeError("tcgen05 pipeline watchdog fired (device-side error flag set)")

    def elbo_step(self, params, grads, ids, eps, mu_out=None, logvar_out=None):
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfga_elbo_step(ctypes.byref(self.desc), _ptr_table(params), _ptr_table(grads), _p(ids), _p(eps),
                                          _p(self.scalars), _p(mu_out), _p(logvar_out), *self._args()))
        return self.scalars

    def capture_elbo_step(self, params, grads, ids, eps):
        self.destroy_graph()
        handle = ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            check(lib.mvae_cfga_elbo_step_graph_create(ctypes.byref(self.desc), _ptr_table(params), _ptr_table(grads), _p(ids),
                                                       _p(eps), _p(self.scalars), _p(None), _p(None), self._ws_ptr,
                                                       self.ws_bytes, ctypes.byref(handle)))
        self._graph = handle
        self._graph_keep = (params, grads, ids, eps)
        return lib.mvae_graph_num_kernel_nodes(handle)

    def launch_graph(self):
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_graph_launch(self._graph, _stream()))
        return self.scalars

    def destroy_graph(self):
        if self._graph is not None:
            lib.mvae_graph_destroy(self._graph)
            self._graph = None
            self._graph_keep = None

    def forward(self, params, ids, eps):
        d = self.desc
        probs = torch.empty(d.batch, d.seq_len, d.charset, dtype=torch.float32, device=self.device)
        mu = torch.empty(d.batch, d.latent, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mu)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfga_forward(ctypes.byref(d), _ptr_table(params), _p(ids), _p(eps), _p(probs), _p(mu), _p(logvar),
                                        *self._args()))
        return probs, mu, logvar

    def backward(self, params, grads, ids, eps, dprobs, dmu, dlogvar, generation=None):
        if generation is not None and generation != self.generation:
            raise _lib.MvaeError(
                "backward() of a forward whose saved activations were overwritten: another forward / decode / fused step ran "
                "on this model between model(x) and loss.backward() (one workspace per module, no forward recomputation)")
        with torch.cuda.device(self.device):
            check(lib.mvae_cfga_backward(ctypes.byref(self.desc), _ptr_table(params), _ptr_table(grads), _p(ids), _p(eps),
                                         _p(dprobs), _p(dmu), _p(dlogvar), *self._args()))

    def decode(self, params, z, want_probs=True):
        d = self.desc
        ids = torch.empty(d.batch, d.seq_len, dtype=torch.uint8, device=self.device)
        probs = torch.empty(d.batch, d.seq_len, d.charset, dtype=torch.float32, device=self.device) if want_probs else None
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfga_decode(ctypes.byref(d), _ptr_table(params), _p(z), _p(ids), _p(probs), *self._args()))
        return ids, probs

    def __del__(self):
        try:
            self.destroy_graph()
        except Exception:
            pass


# ---- the reference's helper modules (models.py:6-77): parameter-free glue, kept so the module tree / state_dict match ----
class Flatten(nn.Module):
    def forward(self, x):
        return x.view(x.size(0), -1)


class Repeat(nn.Module):
    def __init__(self, rep):
        super().__init__()
        self.rep = rep

    def forward(self, x):
        return x.view(x.size(0), 1, -1).repeat(1, self.rep, 1)


class TimeDistributed(nn.Module):
    def __init__(self, module, batch_first=True):
        super().__init__()
        self.module = module
        self.batch_first = batch_first


class SELU(nn.Module):
    def __init__(self, alpha=1.6732632423543772848170429916717, scale=1.0507009873554804934193349852946, inplace=False):
        super().__init__()
        self.scale = scale
        self.elu = nn.ELU(alpha=alpha, inplace=inplace)


def ConvSELU(i, o, kernel_size=3, padding=0, p=0.):
    model = [nn.Conv1d(i, o, kernel_size=kernel_size, padding=padding), SELU(inplace=True)]
    if p > 0.:
        model += [nn.Dropout(p)]
    return nn.Sequential(*model)


class Lambda(nn.Module):
    def __init__(self, i=435, o=292, scale=1E-2):
        super().__init__()
        self.scale = scale
        self.z_mean = nn.Linear(i, o)
        self.z_log_var = nn.Linear(i, o)


class MolEncoder(nn.Module):
    def __init__(self, i=120, o=292, c=35, word_embedding_size=30, h_size=72, num_lstm=3):
        super().__init__()
        self.i = i
        self.embedding = nn.Embedding(num_embeddings=c, embedding_dim=word_embedding_size)
        self.gru = nn.LSTM(word_embedding_size, h_size, num_lstm, batch_first=True)
        self.conv_1 = ConvSELU(i, 120, kernel_size=18)
        self.conv_2 = ConvSELU(120, 64, kernel_size=18)
        self.conv_3 = ConvSELU(64, 64, kernel_size=18)
        self.dense_1 = nn.Sequential(nn.Linear((h_size - (18 * 3) + 3) * 64, 512), SELU(inplace=True))
        self.lmbd = Lambda(512, o)


class MolDecoder(nn.Module):
    def __init__(self, i=292, o=120, c=35, num_gru=4, h_size=1024):
        super().__init__()
        self.latent_input = nn.Sequential(nn.Linear(i, i), SELU(inplace=True))
        self.repeat_vector = Repeat(o)
        self.gru = nn.LSTM(i, h_size, num_gru, batch_first=True)
        self.decoded_mean = TimeDistributed(nn.Sequential(nn.Linear(h_size, c), nn.Softmax()))
        self._owner = None

    def forward(self, x):
        """probabilities (B, o, c) for latents x (B, i) (models.py:161-165)."""
        if self._owner is None:
            raise _lib.MvaeError("MolDecoder runs through its MolecularVAE on this path (model.decoder(z))")
        return self._owner()._decode(x)


class _CfgAFunction(torch.autograd.Function):
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


class MolecularVAE(nn.Module):
    def __init__(self, i=120, o=292, c=35, precision="bf16"):
        super().__init__()
        self.encoder = MolEncoder(i=i, o=o, c=c)
        self.decoder = MolDecoder(i=o, o=i, c=c)
        self.precision = precision
        self._engines = {}
        self.eps_override = None   # tests inject the normal draws here (models.py:92 draws them on the CPU generator)
        self._bind()

    def _bind(self):
        import weakref
        self.decoder._owner = weakref.ref(self)

    def set_submodules(self, encoder=None, decoder=None):
        """Swap in differently sized MolEncoder / MolDecoder instances (their own constructor keywords)."""
        if encoder is not None:
            self.encoder = encoder
        if decoder is not None:
            self.decoder = decoder
        self._engines.clear()
        self._bind()

    # -- plumbing --------------------------------------------------------------------------------------------
    def _cfg(self):
        e, d = self.encoder, self.decoder
        return dict(seq_len=e.i, charset=e.embedding.num_embeddings, embed=e.embedding.embedding_dim,
                    enc_hidden=e.gru.hidden_size, enc_layers=e.gru.num_layers, latent=e.lmbd.z_mean.out_features,
                    dec_hidden=d.gru.hidden_size, dec_layers=d.gru.num_layers, eps_scale=e.lmbd.scale)

    def ordered_params(self):
        """Parameters in the C ABI's order, fetched by attribute path (nn.DataParallel's replicas carry their broadcast
        copies as plain tensor attributes; their named_parameters() is empty -- train_distributed.py:72)."""
        out = []
        for k in cfga_param_order(self.encoder.gru.num_layers, self.decoder.gru.num_layers):
            obj = self
            for part in k.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out

    def engine(self, batch, max_len=None):
        dev = self.encoder.embedding.weight.device
        key = (batch, self.precision)
        # one workspace per (module, device); the dict is shared by nn.DataParallel's per-forward replicas
        with _ENGINE_LOCK:
            slot = self._engines.get(str(dev))
            if slot is None or slot[0] != key:
                cfg = self._cfg()
                eng = CfgAEngine(batch, precision=self.precision, max_len=float(max_len or cfg["seq_len"]), device=dev, **cfg)
                self._engines[str(dev)] = (key, eng)
            else:
                eng = slot[1]
        return eng

    def _eps(self, batch, device):
        if self.eps_override is not None:
            return self.eps_override.to(device=device, dtype=torch.float32).contiguous()
        # models.py:92 draws on the CPU generator and moves the result to the device
        return torch.randn(batch, self.encoder.lmbd.z_mean.out_features).to(device=device, dtype=torch.float32)

    # -- reference API -----------------------------------------------------------------------------------------
    def forward(self, x):
        eng = self.engine(x.shape[0])
        if x.dim() != 2 or x.shape[1] != eng.desc.seq_len:
            raise ValueError(f"expected integer ids of shape (B,{eng.desc.seq_len}), got {tuple(x.shape)}")
        ids = x.to(device=eng.device, dtype=torch.uint8).contiguous()
        eps = self._eps(x.shape[0], ids.device)
        params = [p if p.is_contiguous() else p.contiguous() for p in self.ordered_params()]
        probs, mu, logvar = _CfgAFunction.apply(eng, ids, eps, *params)
        self.encoder.lmbd.mu, self.encoder.lmbd.log_v = mu, logvar   # models.py:90-91 side effects
        return probs, mu, logvar

    @torch.no_grad()
    def _decode(self, z):
        eng = self.engine(z.shape[0])
        _, probs = eng.decode([p.detach() for p in self.ordered_params()], z.detach().float().contiguous(), want_probs=True)
        return probs

    @torch.no_grad()
    def decode_greedy(self, z):
        """argmax ids (B,T) u8 of decoder(z) (train_sample.py:33) without materialising the probabilities."""
        eng = self.engine(z.shape[0])
        ids, _ = eng.decode([p.detach() for p in self.ordered_params()], z.detach().float().contiguous(), want_probs=False)
        return ids

    # -- fused fast path ---------------------------------------------------------------------------------------
    def elbo_step(self, x, eps=None, max_len=None, use_graph=True):
        """Fused forward + loss_function (train.py:31-38) + backward; fills p.grad of every parameter and returns the
        device tensor [loss, max_len*BCE, KL, n_exact_reconstructions]."""
        eng = self.engine(x.shape[0], max_len)
        if max_len is not None and float(max_len) != eng.desc.max_len:
            eng.desc.max_len = float(max_len)
            eng.destroy_graph()
        ids = x.to(device=eng.device, dtype=torch.uint8).contiguous()
        eps = self._eps(x.shape[0], ids.device) if eps is None else eps.to(ids.device, torch.float32).contiguous()
        params = self.ordered_params()
        for p in params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
        if not use_graph:
            return eng.elbo_step([p.data for p in params], [p.grad for p in params], ids, eps)
        key = (tuple(p.data_ptr() for p in params), tuple(p.grad.data_ptr() for p in params))
        if eng._graph is None or getattr(eng, "_graph_key", None) != key:
            eng._ids_static = torch.empty_like(ids)
            eng._eps_static = torch.empty_like(eps)
            eng.capture_elbo_step([p.data for p in params], [p.grad for p in params], eng._ids_static, eng._eps_static)
            eng._graph_key = key
        eng._ids_static.copy_(ids, non_blocking=True)
        eng._eps_static.copy_(eps, non_blocking=True)
        return eng.launch_graph()


max_len = 128  # script-level global read by loss_function (train.py:43)


def loss_function(recon_x, x, mu, logvar):
    """train.py:31-38, verbatim semantics (see models2d.loss_function)."""
    recon_x = recon_x.contiguous().view(-1)
    x = x.contiguous().view(-1)
    xent_loss = max_len * nn.functional.binary_cross_entropy(recon_x, x, reduction="mean")
    kl_loss = -0.5 * torch.mean(1. + mu - logvar ** 2. - torch.exp(mu))
    return xent_loss + kl_loss
