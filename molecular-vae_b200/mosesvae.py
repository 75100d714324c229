"""Drop-in for the reference's mosesvae.py `VAE` (mosesvae.py:27-199) running on the B200 kernels.

Same constructor (`VAE(vocab)` with the vocab duck-type of vocab.py:10-87), same submodules and ModuleList groupings
(`encoder`, `decoder`, `vae` -> identical `state_dict()` keys incl. the aliases, and `model.encoder.parameters()` /
`model.decoder.parameters()` feed separate optimisers as in moses_train_distrib_logp.py:267-268), same
`forward(list[LongTensor]) -> (kl, recon, z, logvar, x_padded, y)`, `string2tensor` / `tensor2string`.
`sample(n_batch, max_len, z, temp) -> (list[str], z)`.  `elbo_step()` is the fused fast path.  In train() mode the
decoder GRU's inter-layer dropout (p = 0.2, mosesvae.py:38,78) is applied with a counter-based mask (the reference's
cuDNN / torch mask stream cannot be reproduced; parity runs inject the same mask into the oracle); eval() disables it.
"""
import ctypes
import os

import torch
from torch import nn

from . import _lib
from ._lib import MosesDesc, check, lib
from .engine import _p, _ptr_table, _stream


def moses_param_order(d_layers=3):
    keys = ["x_emb.weight", "encoder_rnn.weight_ih_l0", "encoder_rnn.weight_hh_l0", "encoder_rnn.bias_ih_l0",
            "encoder_rnn.bias_hh_l0", "q_mu.0.weight", "q_mu.0.bias", "q_mu.2.weight", "q_mu.2.bias",
            "q_logvar.0.weight", "q_logvar.0.bias", "q_logvar.2.weight", "q_logvar.2.bias"]
    for l in range(d_layers):
        keys += [f"decoder_rnn.weight_ih_l{l}", f"decoder_rnn.weight_hh_l{l}", f"decoder_rnn.bias_ih_l{l}",
                 f"decoder_rnn.bias_hh_l{l}"]
    return keys + ["decoder_lat.weight", "decoder_lat.bias", "decoder_fc.weight", "decoder_fc.bias"]


MAX_VOCAB = 256   # token ids travel as u8; logits / one-hot rows are kept in 64-wide tiles (moses.cu make_dims)


def _check_moses_shapes(n_vocab, d_emb, pad, q_d_h, d_d_h):
    """What the kernels assume and torch would otherwise catch as a shape / index error (moses.cu treats x_emb.weight as
    V x V -- the one-hot-initialised table of mosesvae.py:44-50 -- and sizes every table by V)."""
    if d_emb != n_vocab:
        raise ValueError(f"vocab.vectors must be (V, V) one-hot-initialised vectors (vocab.py:87); got width {d_emb} for V={n_vocab}")
    if not 5 <= n_vocab <= MAX_VOCAB:
        raise ValueError(f"vocabulary size {n_vocab} outside the supported range [5, {MAX_VOCAB}]")
    if not 0 <= int(pad) < n_vocab:
        raise ValueError(f"pad id {pad} outside the vocabulary")
    if q_d_h % 64 or d_d_h % 64:
        raise ValueError("GRU hidden sizes must be multiples of 64")


class _MosesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ids, lens, eps, *params):
        ctx.dropout = model._dropout()
        kl, recon, z, logvar, y = model._run(list(params), None, ids, lens, eps, 1.0, 1.0, want_y=True, dropout=ctx.dropout)
        ctx.model, ctx.ids, ctx.lens, ctx.eps = model, ids, lens, eps
        ctx.save_for_backward(*params)
        # z stays differentiable: a property head that consumes it (BindingModel, trainbinding.py:216-217 /
        # moses_train_distrib.py:274) sends its gradient back through dz
        ctx.mark_non_differentiable(logvar, y)
        return kl, recon, z, logvar, y

    @staticmethod
    def backward(ctx, dkl, drecon, dz, dlv, dy):
        params = list(ctx.saved_tensors)
        grads = [torch.empty_like(p) for p in params]
        klw = float(dkl) if dkl is not None else 0.0
        rw = float(drecon) if drecon is not None else 0.0
        dz_ext = None if dz is None else dz.contiguous().float()
        ctx.model._run(params, grads, ctx.ids, ctx.lens, ctx.eps, klw, rw, want_y=False, dropout=ctx.dropout, dz_ext=dz_ext)
        return (None, None, None, None, *grads)


class _BindingFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, z, *params):
        out = model._forward(z, list(params))
        ctx.model, ctx.z, ctx.generation = model, z, model._generation
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.generation != ctx.model._generation:
            raise _lib.MvaeError("BindingModel.backward(): another forward of this module overwrote the saved activations "
                                 "(one workspace per module; call backward before the next forward)")
        params = list(ctx.saved_tensors)
        grads = [torch.empty_like(p) for p in params]
        dz = ctx.model._backward(ctx.z, params, grads, dout.contiguous().float().view(-1), ctx.needs_input_grad[1])
        return (None, dz, *grads)


class BindingModel(nn.Module):
    """Drop-in for mosesvae.BindingModel (mosesvae.py:6-25): same nn.Sequential (hence state_dict keys incl. the BatchNorm
    buffers); forward/backward run in libmvae_b200.so (binding.cu).  A forward is followed by at most one backward."""

    KEYS = ["0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "5.weight", "5.bias", "6.weight", "6.bias",
            "8.weight", "8.bias"]

    def __init__(self, z_size=128):
        super().__init__()
        self.binding_model = nn.Sequential(
            nn.Linear(z_size, 256), nn.BatchNorm1d(256), nn.Tanh(),
            nn.Linear(256, 256), nn.ReLU(),
            nn.Linear(256, 64), nn.BatchNorm1d(64), nn.ReLU(),
            nn.Linear(64, 1))
        self.z_size = z_size
        self._ws = None
        self._generation = 0

    def _desc(self, B):
        bn = self.binding_model[1]
        d = _lib.BindingDesc(B, self.z_size, int(self.training), float(bn.eps), float(bn.momentum))
        need = lib.mvae_binding_workspace_bytes(ctypes.byref(d))
        return d, need

    def _forward(self, z, params):
        if not torch.cuda.is_available():
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        B = z.shape[0]
        d, need = self._desc(B)
        if self._ws is None or self._ws.numel() < need + 256 or self._ws.device != z.device:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=z.device)
        wsp = ctypes.c_void_p(self._ws.data_ptr() + (-self._ws.data_ptr()) % 256)
        bn1, bn2 = self.binding_model[1], self.binding_model[6]
        running = _ptr_table([bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var])
        out = torch.empty(B, dtype=torch.float32, device=z.device)
        self._generation += 1
        with torch.cuda.device(z.device):
            check(lib.mvae_binding_forward(ctypes.byref(d), _ptr_table(params), running, _p(z), _p(out), wsp, need, _stream()))
        if self.training:
            bn1.num_batches_tracked += 1
            bn2.num_batches_tracked += 1
        self._last = (d, wsp, need)
        return out.view(B, 1)

    def _backward(self, z, params, grads, dout, want_dz):
        d, wsp, need = self._last
        dz = torch.empty_like(z) if want_dz else None
        with torch.cuda.device(z.device):
            check(lib.mvae_binding_backward(ctypes.byref(d), _ptr_table(params), _ptr_table(grads), _p(z), _p(dout), _p(dz), wsp,
                                            need, _stream()))
        return dz

    def forward(self, x):
        named = dict(self.binding_model.named_parameters())
        params = [named[k] if named[k].is_contiguous() else named[k].contiguous() for k in self.KEYS]
        return _BindingFunction.apply(self, x.float().contiguous(), *params)


class VAE(nn.Module):
    def __init__(self, vocab, precision="bf16", property_head=False):
        super().__init__()
        q_d_h, q_n_layers, d_n_layers, d_dropout, d_z, d_d_h = 256, 1, 3, 0.2, 160, 512
        self.vocabulary = vocab
        for ss in ("bos", "eos", "unk", "pad"):
            setattr(self, ss, getattr(vocab, ss))
        n_vocab, d_emb = len(vocab), vocab.vectors.size(1)
        _check_moses_shapes(n_vocab, d_emb, self.pad, q_d_h, d_d_h)
        self.x_emb = nn.Embedding(n_vocab, d_emb, self.pad)
        self.x_emb.weight.data.copy_(vocab.vectors)
        self.encoder_rnn = nn.GRU(d_emb, q_d_h, num_layers=q_n_layers, batch_first=True, dropout=0, bidirectional=False)
        self.q_mu = nn.Sequential(nn.Linear(q_d_h, 256), nn.ReLU(), nn.Linear(256, d_z))
        self.q_logvar = nn.Sequential(nn.Linear(q_d_h, 256), nn.ReLU(), nn.Linear(256, d_z))
        self.decoder_rnn = nn.GRU(d_emb + d_z, d_d_h, num_layers=d_n_layers, batch_first=True, dropout=d_dropout)
        self.decoder_lat = nn.Linear(d_z, d_d_h)
        self.decoder_fc = nn.Linear(d_d_h, n_vocab)
        self.encoder = nn.ModuleList([self.x_emb, self.encoder_rnn, self.q_mu, self.q_logvar])
        self.decoder = nn.ModuleList([self.decoder_rnn, self.decoder_lat, self.decoder_fc])
        self.vae = nn.ModuleList([self.x_emb, self.encoder, self.decoder])
        self.precision = precision
        self.cfg = dict(vocab=n_vocab, d_z=d_z, q_hidden=q_d_h, d_hidden=d_d_h, d_layers=d_n_layers, mlp_hidden=256)
        self._keys = moses_param_order(d_n_layers)
        self._ws = None
        self.eps_override = None
        self.dropout_seed_override = None   # tests pin the counter-based dropout mask here
        if property_head:
            self.attach_property_head()

    def attach_property_head(self, head=None):
        """`model.binding_model`: the BindingModel MLP on z that the historical VAE carried (mosesanalyize.py:177 optimises
        `model.binding_model.parameters()`; moses_train_distrib.py:274 calls `model(input_batch, binding)`).  It is not a
        member of the encoder / decoder / vae ModuleLists, so those optimiser groups stay as in mosesvae.py:90-105."""
        self.binding_model = head if head is not None else BindingModel(self.cfg["d_z"])
        return self.binding_model

    @property
    def device(self):
        return next(self.parameters()).device

    def string2tensor(self, string, device="model"):
        ids = self.vocabulary.string2ids(string, add_bos=True, add_eos=True)
        return torch.tensor(ids, dtype=torch.long, device=self.device if device == "model" else device)

    def tensor2string(self, tensor):
        return self.vocabulary.ids2string(tensor.tolist(), rem_bos=True, rem_eos=True)

    def ordered_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in self._keys]

    # -- plumbing ------------------------------------------------------------------------------------
    def _pack(self, x):
        """list of id tensors (sorted by length desc, as pack_sequence requires) -> u8 (B,T) padded, int32 lengths"""
        lens = [int(t.numel()) for t in x]
        if any(lens[i] < lens[i + 1] for i in range(len(lens) - 1)):
            raise RuntimeError("sequences must be sorted by length in decreasing order (pack_sequence contract)")
        if min(lens) < 2:
            raise RuntimeError("every sequence needs at least <bos> and <eos>")
        dev = self.device
        lens_h = torch.tensor(lens, dtype=torch.int32)
        # one concatenation + one scatter instead of pad_sequence's per-sequence copies (4096 tiny kernels at B = 4096)
        B, T = len(x), lens[0]
        flat = torch.cat([t.reshape(-1) for t in x]).to(device=dev, dtype=torch.long)
        row = torch.repeat_interleave(torch.arange(B), lens_h.long())
        col = torch.arange(int(lens_h.sum())) - torch.repeat_interleave(torch.cumsum(lens_h.long(), 0) - lens_h.long(), lens_h.long())
        padded = torch.full((B, T), int(self.pad), dtype=torch.long, device=dev)
        padded.view(-1).index_copy_(0, (row * T + col).to(dev), flat)
        # token ids index device tables: an id outside the vocabulary is an error the way it is for nn.Embedding.  The
        # check is asynchronous (no host sync on the training path): ids are clamped for the kernels and the flag is
        # raised by check_device_error() / the next forward().
        V = self.cfg["vocab"]
        bad = ((padded < 0) | (padded >= V)).any()
        self._bad_ids = bad if getattr(self, "_bad_ids", None) is None else (self._bad_ids | bad)
        padded_c = padded.clamp(0, V - 1)
        lens_d = lens_h.to(dev)
        lens_d._host_copy = lens_h
        return padded, padded_c.to(torch.uint8).contiguous(), lens_d

    def _dropout(self):
        """(p, seed) of the train-mode dropout between decoder layers (mosesvae.py:38,78); p = 0 in eval mode."""
        p = float(getattr(self.decoder_rnn, "dropout", 0.0)) if self.training else 0.0
        if p <= 0.0:
            return 0.0, 0
        if self.dropout_seed_override is not None:
            return p, int(self.dropout_seed_override)
        return p, int(torch.randint(0, 2 ** 31 - 1, (1,)).item())

    def _desc(self, B, T, kl_weight=1.0, recon_weight=1.0, dropout=(0.0, 0)):
        c = self.cfg
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[self.precision]
        return MosesDesc(B, T, c["vocab"], c["d_z"], c["q_hidden"], c["d_hidden"], c["d_layers"], c["mlp_hidden"],
                         int(self.pad), prec, float(kl_weight), float(recon_weight), int(c.get("q_bidir", 0)),
                         int(c.get("q_linear_heads", 0)), float(dropout[0]), int(dropout[1]))

    def _run(self, params, grads, ids, lens, eps, kl_weight, recon_weight, want_y, dropout=(0.0, 0), dz_ext=None, phase=-1,
             joint=None):
        """One call into the C ABI.  joint = (binding_model, target fp32 (B), weight, binding_grads or None): the VAE step with
        the property head inside (mvae_moses_joint_step); otherwise mvae_moses_step_ex."""
        if not torch.cuda.is_available():
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        B, T = ids.shape
        c = self.cfg
        d = self._desc(B, T, kl_weight, recon_weight, dropout)
        need = lib.mvae_moses_workspace_bytes(ctypes.byref(d))
        if need == 0:
            raise ValueError("invalid MOSES VAE description")
        dev = ids.device
        if self._ws is None or self._ws.numel() < need + 256 or self._ws.device != dev:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        wsp = ctypes.c_void_p(self._ws.data_ptr() + (-self._ws.data_ptr()) % 256)
        # later phases of a phased step keep writing the scalars / z of phase 0
        keep = getattr(self, "_phase_outputs", None) if phase > 0 else None
        out = keep[0] if keep else torch.zeros(4, dtype=torch.float32, device=dev)
        z = keep[1] if keep else torch.empty(B, c["d_z"], dtype=torch.float32, device=dev)
        lv = keep[2] if keep else torch.empty_like(z)
        if phase == 0:
            self._phase_outputs = (out, z, lv)
        y = torch.empty(B, T, c["vocab"], dtype=torch.float32, device=dev) if want_y else None
        P = _ptr_table(params)
        G = _ptr_table(grads) if grads is not None else None
        lens_host = getattr(lens, "_host_copy", None)      # set by _pack: enables packed-sequence batch sizes per step
        with torch.cuda.device(dev):
            if joint is None:
                check(lib.mvae_moses_step_ex(ctypes.byref(d), P, G, _p(ids), _p(lens), _p(lens_host), _p(eps), _p(dz_ext), _p(out),
                                             _p(z), _p(lv), _p(y), wsp, need, int(phase), _stream()))
            else:
                head, target, weight, bgrads = joint
                bd, bneed = head._desc(B)
                if head._ws is None or head._ws.numel() < bneed + 256 or head._ws.device != dev:
                    head._ws = torch.empty(bneed + 256, dtype=torch.uint8, device=dev)
                bwsp = ctypes.c_void_p(head._ws.data_ptr() + (-head._ws.data_ptr()) % 256)
                xneed = lib.mvae_moses_joint_extra_bytes(ctypes.byref(d))
                if getattr(self, "_joint_extra", None) is None or self._joint_extra.numel() < xneed + 256 or self._joint_extra.device != dev:
                    self._joint_extra = torch.empty(xneed + 256, dtype=torch.uint8, device=dev)
                    self._joint_loss = torch.zeros(1, dtype=torch.float32, device=dev)
                xp = ctypes.c_void_p(self._joint_extra.data_ptr() + (-self._joint_extra.data_ptr()) % 256)
                bn1, bn2 = head.binding_model[1], head.binding_model[6]
                named = dict(head.binding_model.named_parameters())
                bparams = [named[k].data for k in head.KEYS]
                running = _ptr_table([bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var])
                check(lib.mvae_moses_joint_step(ctypes.byref(d), P, G, _p(ids), _p(lens), _p(lens_host), _p(eps), ctypes.byref(bd),
                                                _ptr_table(bparams), _ptr_table(bgrads) if bgrads is not None else None, running,
                                                _p(target), float(weight), _p(out), _p(self._joint_loss), _p(z), wsp, need, bwsp,
                                                bneed, xp, xneed, int(phase), _stream()))
                head._generation += 1
                self.last_binding_loss = self._joint_loss
                if head.training and phase <= 0:
                    bn1.num_batches_tracked += 1
                    bn2.num_batches_tracked += 1
        self._last_desc = (d, wsp, need)
        self._last_scalars = out
        return out[1], out[2], z, lv, y

    def check_device_error(self):
        bad = getattr(self, "_bad_ids", None)
        if bad is not None:
            self._bad_ids = None
            if bool(bad.item()):
                raise IndexError("token id outside the vocabulary (as nn.Embedding would raise, mosesvae.py:150)")
        d, wsp, need = self._last_desc
        flag = ctypes.c_int(0)
        check(lib.mvae_moses_read_error(ctypes.byref(d), wsp, need, ctypes.byref(flag), _stream()))
        if flag.value:
            raise _lib.MvaeError("tcgen05 pipeline watchdog fired (device-side error flag set)")

    def _eps(self, B, dev):
        if self.eps_override is not None:
            return self.eps_override.to(device=dev, dtype=torch.float32).contiguous()
        return torch.randn(B, self.cfg["d_z"], device=dev, dtype=torch.float32)

    # -- reference API (mosesvae.py:126-140) ------------------------------------------------------------
    def forward(self, x, binding=None):
        """mosesvae.py:126-140: (kl, recon, z, logvar, x_padded, y).  With `binding` (the property target, (B,) or (B,1)) the
        historical signature of moses_train_distrib.py:274 / trainbinding.py:216: (kl, recon, binding_loss, z), where
        binding_loss = mse(self.binding_model(z), binding) and z carries the autograd graph back into the encoder."""
        x_pad, ids, lens = self._pack(x)
        eps = self._eps(ids.shape[0], ids.device)
        params = [p if p.is_contiguous() else p.contiguous() for p in self.ordered_params()]
        kl, recon, z, logvar, y = _MosesFunction.apply(self, ids, lens, eps, *params)
        if binding is None:
            return kl, recon, z, logvar, x_pad, y
        if getattr(self, "binding_model", None) is None:
            raise RuntimeError("forward(x, binding) needs the property head: VAE(vocab, property_head=True) or attach_property_head()")
        pred = self.binding_model(z)
        binding_loss = nn.functional.mse_loss(pred, binding.to(pred.device, torch.float32).view(-1, 1))
        return kl, recon, binding_loss, z

    def elbo_step(self, x, kl_weight=1.0, recon_weight=1.0, eps=None, binding=None, binding_weight=1.0):
        """Fused forward + backward of kl_weight*kl + recon_weight*recon (+ binding_weight*mse(binding_model(z), binding) when
        a property target is given): fills p.grad (also of the property head) and returns the device tensor
        [loss, kl, recon, n_targets]; the head's loss term is `self.last_binding_loss` (1-element device tensor)."""
        _, ids, lens = self._pack(x)
        eps = self._eps(ids.shape[0], ids.device) if eps is None else eps.to(ids.device, torch.float32).contiguous()
        params = self.ordered_params()
        for p in params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
        joint = None
        if binding is not None:
            head = getattr(self, "binding_model", None)
            if head is None:
                raise RuntimeError("elbo_step(binding=...) needs the property head: VAE(vocab, property_head=True)")
            named = dict(head.binding_model.named_parameters())
            for k in head.KEYS:
                if named[k].grad is None:
                    named[k].grad = torch.empty_like(named[k])
            joint = (head, binding.to(ids.device, torch.float32).contiguous().view(-1), binding_weight, [named[k].grad for k in head.KEYS])
        self._run([p.data for p in params], [p.grad for p in params], ids, lens, eps, kl_weight, recon_weight, False,
                  dropout=self._dropout(), joint=joint)
        return self._last_scalars

    def sample_z_prior(self, n_batch):
        """z ~ N(0, I).  The shipped mosesvae.py:201-211 returns zeros through an undefined `self.d_z`; the working
        prior is mosesfile.py:165-166 (randn), which is what this returns."""
        return torch.randn(n_batch, self.cfg["d_z"], device=self.device)

    @torch.no_grad()
    def sample_ids(self, n_batch, max_len=100, z=None, temp=1.0, greedy=False, seed=None, use_graph=True):
        """Device-side result of sample(): (ids u8 (B,max_len), lengths int32 (B), z).  With use_graph the whole decode
        loop (mosesvae.py:239-251, 99 steps) is one CUDA-graph replay over fixed buffers: z and the generator seed are
        written into them before each replay, the weights are re-read from the parameters every time."""
        if z is None:
            z = self.sample_z_prior(n_batch)
        z = z.to(self.device, torch.float32).contiguous()
        B, c = z.shape[0], self.cfg
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[self.precision]
        d = MosesDesc(B, int(max_len), c["vocab"], c["d_z"], c["q_hidden"], c["d_hidden"], c["d_layers"], c["mlp_hidden"],
                      int(self.pad), prec, 1.0, 1.0, int(c.get("q_bidir", 0)), int(c.get("q_linear_heads", 0)), 0.0, 0)
        need = lib.mvae_moses_workspace_bytes(ctypes.byref(d))
        dev = z.device
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        params = [p.detach() for p in self.ordered_params()]
        mode = 0 if greedy else 1
        if use_graph:
            key = (B, int(max_len), mode, float(temp), self.precision, dev, os.environ.get("MVAE_SAMPLE_FUSED", ""),
                   os.environ.get("MVAE_SAMPLE_PERSISTENT", ""),
                   tuple(p.data_ptr() for p in params))
            g = getattr(self, "_sample_graph", None)
            if g is None or g["key"] != key:
                self.destroy_sample_graph()
                ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
                g = {"key": key, "ws": ws, "z": torch.empty_like(z), "seed": torch.zeros(1, dtype=torch.int64, device=dev),
                     "ids": torch.empty(B, int(max_len), dtype=torch.uint8, device=dev),
                     "lens": torch.empty(B, dtype=torch.int32, device=dev), "desc": d, "params": params,
                     "ptrs": _ptr_table(params)}
                g["wsp"] = ctypes.c_void_p(ws.data_ptr() + (-ws.data_ptr()) % 256)
                handle = ctypes.c_void_p()
                with torch.cuda.device(dev):
                    torch.cuda.current_stream().synchronize()
                    check(lib.mvae_moses_sample_graph_create(ctypes.byref(d), g["ptrs"], _p(g["z"]), int(self.bos),
                                                             int(self.eos), mode, float(temp), _p(g["seed"]), _p(g["ids"]),
                                                             _p(g["lens"]), g["wsp"], need, ctypes.byref(handle)))
                g["handle"] = handle
                self._sample_graph = g
            g["z"].copy_(z)
            g["seed"].fill_(seed)
            with torch.cuda.device(dev):
                check(lib.mvae_graph_launch(g["handle"], _stream()))
            self._last_desc = (g["desc"], g["wsp"], need)
            return g["ids"].clone(), g["lens"].clone(), z
        if self._ws is None or self._ws.numel() < need + 256 or self._ws.device != dev:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        wsp = ctypes.c_void_p(self._ws.data_ptr() + (-self._ws.data_ptr()) % 256)
        ids = torch.empty(B, int(max_len), dtype=torch.uint8, device=dev)
        lens = torch.empty(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mvae_moses_sample(ctypes.byref(d), _ptr_table(params), _p(z), int(self.bos), int(self.eos),
                                        mode, float(temp), ctypes.c_ulonglong(seed), None, _p(ids), _p(lens), wsp,
                                        need, _stream()))
        self._last_desc = (d, wsp, need)
        return ids, lens, z

    def sample_many(self, n_total, n_batch=8192, max_len=100, temp=1.0, greedy=False, seed=None, table=None):
        """hugesample.py:25-40 as a generator: decodes n_total latents from the prior in batches of n_batch and yields one
        list[str] per batch.  The GPU decodes batch k+1 (one CUDA-graph replay) while the host turns batch k's bytes into
        Python strings; per batch one device-side text assembly and one device -> pinned-host copy.  `table`: a
        text.TokenTable (e.g. the "[sym]" join of hugesample.py:34); default = the vocabulary's ids2string."""
        if table is None:
            table = getattr(self, "_token_table", None)
            if table is None or table.table.device != self.device:
                from .text import TokenTable
                table = self._token_table = TokenTable.from_vocab(self.vocabulary, self.device)
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        pending, done = None, 0
        while done < n_total:
            b = n_batch   # fixed batch: one captured graph; the last batch is trimmed on the host
            s = int(torch.randint(0, 2 ** 62, (1,), generator=gen).item())
            ids, lens, _ = self.sample_ids(b, max_len=max_len, temp=temp, greedy=greedy, seed=s)
            nxt = table.to_strings_async(ids, lens)
            if pending is not None:
                yield pending[0].result()[:pending[1]]
            pending = (nxt, min(b, n_total - done))
            done += b
        if pending is not None:
            yield pending[0].result()[:pending[1]]

    def destroy_sample_graph(self):
        g = getattr(self, "_sample_graph", None)
        if g is not None:
            lib.mvae_graph_destroy(g["handle"])
            self._sample_graph = None

    def sample(self, n_batch, max_len=100, z=None, temp=1.0, greedy=False, seed=None):
        """mosesvae.py:214-262: returns (list[str], z).  Multinomial draws by default (the reference behaviour),
        greedy=True selects argmax decoding (the bit-exact parity mode)."""
        ids, lens, z = self.sample_ids(n_batch, max_len, z, temp, greedy, seed)
        # mosesvae.py:258-262 slices every row and calls tensor2string on it (B .tolist() round trips); here the strings
        # are assembled on the device and come back in one copy
        tab = getattr(self, "_token_table", None)
        if tab is None or tab.table.device != ids.device:
            from .text import TokenTable
            tab = self._token_table = TokenTable.from_vocab(self.vocabulary, ids.device)
        return tab.to_strings(ids, lens), z
