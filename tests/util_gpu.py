"""Shared helpers for the GPU parity tests (oracle = oracle/vae_oracle.py, float64)."""
import numpy as np
import torch

from oracle import vae_oracle as vo


def load_pkg():
    import molecular_vae_b200 as m
    return m


def make_case(param_seed, batch_seed, B, Z, H, L, dtype=np.float32):
    P = vo.make_params(param_seed, dtype=np.float32, latent=Z, hidden=H, layers=L)
    ids, onehot, eps = vo.make_batch(batch_seed, B, latent=Z, dtype=np.float32)
    return P, ids, onehot, eps


def oracle_step(P, onehot, eps, L, train=True, max_len=120, need_grads=True):
    P64 = {k: v.astype(np.float64) for k, v in P.items()}
    return vo.config_b_step(P64, onehot.astype(np.float64), eps.astype(np.float64), max_len=max_len, train=train,
                            need_grads=need_grads, layers=L)


def build_model(m, P, Z, H, L, precision):
    model = m.VAE(latent=Z, hidden=H, layers=L, precision=precision)
    sd = {k: torch.from_numpy(v) for k, v in P.items()}
    model.load_state_dict(sd, strict=True)
    return model.cuda()


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / (np.sqrt((b ** 2).sum()) + 1e-300))


def grads_of(model):
    return {k: p.grad.detach().cpu().numpy() for k, p in model.named_parameters()}
