"""Host-side driver of the Config-B ELBO step: owns the workspace, builds the pointer tables and calls the
C ABI on torch's current CUDA stream.  PyTorch is used for device memory and streams only."""
import ctypes

import torch

from . import _lib
from ._lib import CfgBDesc, check, lib

PARAM_ORDER_HEAD = ["conv1d1.weight", "conv1d1.bias", "conv1d2.weight", "conv1d2.bias", "conv1d3.weight",
                    "conv1d3.bias", "fc0.weight", "fc0.bias", "fc11.weight", "fc11.bias", "fc12.weight", "fc12.bias",
                    "fc2.weight", "fc2.bias"]


def param_order(layers):
    """state_dict keys in the order of the C ABI's pointer table (include/mvae_b200.h)."""
    keys = list(PARAM_ORDER_HEAD)
    for l in range(layers):
        keys += [f"gru.weight_ih_l{l}", f"gru.weight_hh_l{l}", f"gru.bias_ih_l{l}", f"gru.bias_hh_l{l}"]
    return keys + ["fc3.weight", "fc3.bias"]


def _ptr_table(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("parameters / gradients must be contiguous fp32 CUDA tensors")
        arr[i] = t.data_ptr()
    return arr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class CfgBEngine:
    """One instance per (shape, precision).  Holds the HBM workspace (activations for BPTT, padded weights)."""

    def __init__(self, batch, seq_len=120, charset=35, latent=292, hidden=501, layers=3, fc0=435,
                 precision="bf16", train=True, max_len=120.0, eps_scale=1.0, device=None):
        if not torch.cuda.is_available():
            raise _lib.MvaeError("molecular-vae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        prec = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16}[precision]
        self.precision = precision
        self.desc = CfgBDesc(batch, seq_len, charset, latent, hidden, layers, fc0, prec, int(bool(train)),
                             float(max_len), float(eps_scale))
        self.ws_bytes = lib.mvae_cfgb_workspace_bytes(ctypes.byref(self.desc))
        if self.ws_bytes == 0:
            raise ValueError("invalid Config-B description")
        with torch.cuda.device(self.device):
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
        off = (-self.ws.data_ptr()) % 256
        self._ws_ptr = ctypes.c_void_p(self.ws.data_ptr() + off)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._graph = None
        self._graph_keep = None
        self._phase_graphs = []
        self._phase_keep = None
        # every call that rewrites the workspace bumps this; an autograd backward checks that the activations it is about
        # to read are still the ones its own forward left there (one shared workspace per engine)
        self.generation = 0

    # -- helpers -------------------------------------------------------------------------------
    @property
    def batch(self):
        return self.desc.batch

    def set_train(self, train):
        self.desc.train = int(bool(train))

    def to_ids(self, x):
        """(B,T) integer ids or (B,T,C) float one-hot -> contiguous u8 ids on the device."""
        d = self.desc
        if x.dim() == 2:
            return x.to(device=self.device, dtype=torch.uint8).contiguous()
        if x.dim() != 3 or x.shape[1] != d.seq_len or x.shape[2] != d.charset:
            raise ValueError(f"expected one-hot of shape (B,{d.seq_len},{d.charset}), got {tuple(x.shape)}")
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        ids = torch.empty(x.shape[:2], dtype=torch.uint8, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.mvae_onehot_to_ids(_p(x), x.shape[0] * x.shape[1], d.charset, _p(ids), _p(flag), _stream()))
        if int(flag.item()) != 0:
            raise ValueError("input is not exactly one-hot; the B200 path consumes character ids")
        return ids

    def check_device_error(self):
        flag = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_read_error(ctypes.byref(self.desc), self._ws_ptr, self.ws_bytes, ctypes.byref(flag),
                                           _stream()))
        if flag.value:
            raise _lib.MvaeError("tcgen05 pipeline watchdog fired (device-side error flag set)")

    # -- fused step ----------------------------------------------------------------------------
    def elbo_step(self, params, grads, ids, eps, mu_out=None, logvar_out=None):
        """Fused forward + loss + backward.  Returns the device tensor [loss, max_len*bce, kl, n_exact]."""
        P, G = _ptr_table(params), _ptr_table(grads)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_elbo_step(ctypes.byref(self.desc), P, G, _p(ids), _p(eps), _p(self.scalars),
                                          _p(mu_out), _p(logvar_out), self._ws_ptr, self.ws_bytes, _stream()))
        return self.scalars

    def capture_elbo_step(self, params, grads, ids, eps, mu_out=None, logvar_out=None):
        """Capture the fused step into a CUDA graph over FIXED buffers; replay with launch_graph()."""
        self.destroy_graph()
        P, G = _ptr_table(params), _ptr_table(grads)
        handle = ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            check(lib.mvae_cfgb_elbo_step_graph_create(ctypes.byref(self.desc), P, G, _p(ids), _p(eps),
                                                       _p(self.scalars), _p(mu_out), _p(logvar_out), self._ws_ptr,
                                                       self.ws_bytes, ctypes.byref(handle)))
        self._graph = handle
        self._graph_keep = (params, grads, ids, eps, mu_out, logvar_out)
        return lib.mvae_graph_num_kernel_nodes(handle)

    def capture_elbo_step_phases(self, params, grads, ids, eps):
        """Capture the step as `layers` CUDA graphs (mvae_cfgb_elbo_step_phase): after phase p the gradient bucket
        `ddp.phase_buckets(...)[p]` is final, so its all-reduce can overlap the next phase."""
        self.destroy_phase_graphs()
        P, G = _ptr_table(params), _ptr_table(grads)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for phase in range(self.desc.layers):
                handle = ctypes.c_void_p(0)
                check(lib.mvae_cfgb_elbo_step_phase_graph_create(ctypes.byref(self.desc), P, G, _p(ids), _p(eps),
                                                                 _p(self.scalars), _p(None), _p(None), self._ws_ptr,
                                                                 self.ws_bytes, phase, ctypes.byref(handle)))
                self._phase_graphs.append(handle)
        self._phase_keep = (params, grads, ids, eps)
        return [lib.mvae_graph_num_kernel_nodes(h) for h in self._phase_graphs]

    def elbo_step_phase(self, params, grads, ids, eps, phase):
        """One phase of the fused step as a direct launch on the current stream (mvae_cfgb_elbo_step_phase): used inside a
        torch.cuda.graph capture that also holds the gradient all-reduces (ddp.GraphedDataParallelStep)."""
        P, G = _ptr_table(params), _ptr_table(grads)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_elbo_step_phase(ctypes.byref(self.desc), P, G, _p(ids), _p(eps), _p(self.scalars), _p(None),
                                                _p(None), self._ws_ptr, self.ws_bytes, int(phase), _stream()))
        return self.scalars

    def launch_phase(self, phase):
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_graph_launch(self._phase_graphs[phase], _stream()))
        return self.scalars

    def destroy_phase_graphs(self):
        for h in getattr(self, "_phase_graphs", []):
            lib.mvae_graph_destroy(h)
        self._phase_graphs = []
        self._phase_keep = None

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

    # -- drop-in forward / backward -------------------------------------------------------------
    def forward(self, params, ids, eps):
        d = self.desc
        probs = torch.empty(d.batch, d.seq_len, d.charset, dtype=torch.float32, device=self.device)
        mu = torch.empty(d.batch, d.latent, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mu)
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_forward(ctypes.byref(d), _ptr_table(params), _p(ids), _p(eps), _p(probs), _p(mu),
                                        _p(logvar), self._ws_ptr, self.ws_bytes, _stream()))
        return probs, mu, logvar

    def backward(self, params, grads, ids, eps, dprobs, dmu, dlogvar, generation=None):
        if generation is not None and generation != self.generation:
            raise _lib.MvaeError(
                "backward() of a forward whose saved activations were overwritten: another forward / decode / fused step ran "
                "on this model between model(x) and loss.backward().  Call backward first, or use a second model instance "
                "(the engine keeps ONE workspace per module; it does not recompute the forward)")
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_backward(ctypes.byref(self.desc), _ptr_table(params), _ptr_table(grads), _p(ids),
                                         _p(eps), _p(dprobs), _p(dmu), _p(dlogvar), self._ws_ptr, self.ws_bytes,
                                         _stream()))

    def decode_greedy(self, params, z, want_probs=False):
        d = self.desc
        ids = torch.empty(d.batch, d.seq_len, dtype=torch.uint8, device=self.device)
        probs = torch.empty(d.batch, d.seq_len, d.charset, dtype=torch.float32, device=self.device) if want_probs else None
        self.generation += 1
        with torch.cuda.device(self.device):
            check(lib.mvae_cfgb_decode_greedy(ctypes.byref(d), _ptr_table(params), _p(z), _p(ids), _p(probs),
                                              self._ws_ptr, self.ws_bytes, _stream()))
        return ids, probs

    def __del__(self):
        try:
            self.destroy_graph()
            self.destroy_phase_graphs()
        except Exception:
            pass
