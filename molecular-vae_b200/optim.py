"""Fused optimiser step over flat buffers: clip_grad_norm + Adam / SGD-momentum (train.py:81,102-104,
train_distributed.py:73,91-94) in three kernel launches and no host synchronisation."""
import ctypes

import torch

from ._lib import check, lib
from .ddp import FlatGradBuffer


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams:
    """Re-homes every parameter into ONE flat fp32 buffer (p.data becomes a view) next to a FlatGradBuffer."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        self.flat = torch.empty(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            off += n
        self.grads = FlatGradBuffer(self.params)


class FusedOptimizer:
    def __init__(self, flat_params: FlatParams, kind="adam", lr=8e-4, betas=(0.9, 0.999), eps=1e-8, momentum=0.85,
                 weight_decay=0.0, max_norm=None):
        self.fp, self.kind, self.lr, self.betas, self.eps = flat_params, kind, lr, betas, eps
        self.momentum, self.weight_decay, self.max_norm = momentum, weight_decay, max_norm
        n, dev = flat_params.flat.numel(), flat_params.flat.device
        self.state1 = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state2 = torch.zeros(n, dtype=torch.float32, device=dev) if kind == "adam" else None
        self.scratch = torch.zeros(4, dtype=torch.float32, device=dev)   # 16 bytes: double sumsq + float coef
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.steps = 0

    # -- torch.optim surface the reference's loops touch (train.py:81-83,104,170-177: ReduceLROnPlateau reads / writes
    #    param_groups[0]['lr'], the checkpoint stores optimizer.state_dict()) ------------------------------------------
    @property
    def param_groups(self):
        outer = self

        class _Group(dict):
            def __setitem__(g, k, v):
                dict.__setitem__(g, k, v)
                if k == "lr":
                    outer.lr = float(v)

        g = _Group(params=self.fp.params, lr=self.lr, weight_decay=self.weight_decay)
        if self.kind == "adam":
            g.update(betas=self.betas, eps=self.eps)
        else:
            g.update(momentum=self.momentum)
        return [g]

    def zero_grad(self, set_to_none=False):
        self.fp.grads.flat.zero_()

    def _slices(self):
        off = 0
        for i, p in enumerate(self.fp.params):
            yield i, p, off, off + p.numel()
            off += p.numel()

    def state_dict(self):
        """torch.optim.Adam / SGD layout: per-parameter `exp_avg`, `exp_avg_sq`, `step` (Adam) or `momentum_buffer` (SGD)
        cut out of the flat state buffers, and one param group -- what train.py:170-177 stores as optimizer_state_dict."""
        state = {}
        if self.steps > 0:
            for i, p, lo, hi in self._slices():
                if self.kind == "adam":
                    state[i] = {"step": torch.tensor(float(self.steps)), "exp_avg": self.state1[lo:hi].view_as(p).clone(),
                                "exp_avg_sq": self.state2[lo:hi].view_as(p).clone()}
                else:
                    state[i] = {"momentum_buffer": self.state1[lo:hi].view_as(p).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.fp.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        self.lr = float(group["lr"])
        self.weight_decay = float(group.get("weight_decay", self.weight_decay))
        if self.kind == "adam":
            self.betas, self.eps = tuple(group.get("betas", self.betas)), float(group.get("eps", self.eps))
        else:
            self.momentum = float(group.get("momentum", self.momentum))
        self.steps = 0
        self.state1.zero_()
        if self.state2 is not None:
            self.state2.zero_()
        for i, p, lo, hi in self._slices():
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            if self.kind == "adam":
                self.state1[lo:hi].copy_(st["exp_avg"].reshape(-1))
                self.state2[lo:hi].copy_(st["exp_avg_sq"].reshape(-1))
                self.steps = max(self.steps, int(float(st["step"])))
            else:
                self.state1[lo:hi].copy_(st["momentum_buffer"].reshape(-1))
                self.steps = max(self.steps, 1)

    def step(self):
        fp, n = self.fp, self.fp.flat.numel()
        coef = ctypes.c_void_p(0)
        with torch.cuda.device(fp.flat.device):
            if self.max_norm is not None:
                check(lib.mvae_clip_grad_norm(_p(fp.grads.flat), n, float(self.max_norm), _p(self.scratch),
                                              _p(self.norm), 0, _stream()))
                coef = ctypes.c_void_p(self.scratch.data_ptr() + 8)
            self.steps += 1
            if self.kind == "adam":
                check(lib.mvae_adam_step(_p(fp.flat), _p(fp.grads.flat), _p(self.state1), _p(self.state2), n,
                                         float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                         float(self.weight_decay), self.steps, coef, _stream()))
            else:
                check(lib.mvae_sgd_momentum_step(_p(fp.flat), _p(fp.grads.flat), _p(self.state1), n, float(self.lr),
                                                 float(self.momentum), float(self.weight_decay),
                                                 int(self.steps == 1), coef, _stream()))
