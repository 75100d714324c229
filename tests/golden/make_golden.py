"""Generate golden vectors from the UNMODIFIED reference modules (run in the build
container only: /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Imports /root/reference/models2d.py (Config-B layer stack, latent widened 2 -> Z the
way SURVEY.md 8c describes) and AST-extracts loss_function from train.py:31-38 (the
script itself imports comet_ml and reads absolute paths).  Weights / inputs / eps come
from oracle.vae_oracle.make_params / make_batch (numpy PCG64, seed-addressed) so the
fixtures only need to hold OUTPUTS.  eps is injected by patching torch.randn_like for
the duration of the reference forward (models2d.py:34) -- the reference is not edited.
Everything is evaluated in float64 (the tolerance anchor) and float32 (what the
reference actually runs in).
"""
import ast
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from oracle import vae_oracle as vo  # noqa: E402


def load_loss_function(max_len):
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "loss_function"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "nn": torch.nn, "max_len": max_len}
    exec(compile(mod, "train.py[loss_function]", "exec"), ns)
    return ns["loss_function"]


def build_reference_model(P, latent, hidden, layers, dtype):
    import models2d  # the reference module
    m = models2d.VAE()
    nn = torch.nn
    m.fc11 = nn.Linear(435, latent)
    m.fc12 = nn.Linear(435, latent)
    m.fc2 = nn.Linear(latent, latent)
    m.gru = nn.GRU(latent, hidden, layers, batch_first=True)
    m.fc3 = nn.Linear(hidden, 35)
    m = m.to(dtype)
    sd = {k: torch.from_numpy(np.asarray(v)).to(dtype) for k, v in P.items()}
    m.load_state_dict(sd, strict=True)
    return m


def run_reference(P, onehot, eps, latent, hidden, layers, dtype, train=True, max_len=120):
    m = build_reference_model(P, latent, hidden, layers, dtype)
    m.train(train)
    x = torch.from_numpy(onehot).to(dtype)
    e = torch.from_numpy(eps).to(dtype)
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: e
    try:
        probs, mu, logvar = m(x)
    finally:
        torch.randn_like = orig
    lf = load_loss_function(max_len)
    loss = lf(probs, x, mu, logvar)
    # terms, restated only to SPLIT the scalar the reference returns
    bce = max_len * torch.nn.functional.binary_cross_entropy(probs.reshape(-1), x.reshape(-1))
    kl = loss - bce
    loss.backward()
    grads = {k: p.grad.detach().numpy().astype(np.float64) for k, p in m.named_parameters()}
    return dict(loss=float(loss), bce=float(bce), kl=float(kl),
                probs=probs.detach().numpy(), mu=mu.detach().numpy(), logvar=logvar.detach().numpy(),
                grads=grads)


def sample_idx(n, k=64, seed=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.sort(rng.choice(n, size=min(k, n), replace=False))


def pack(res, full_grads):
    out = dict(loss=res["loss"], bce=res["bce"], kl=res["kl"],
               probs=res["probs"].astype(np.float64), mu=res["mu"].astype(np.float64),
               logvar=res["logvar"].astype(np.float64))
    for k, g in res["grads"].items():
        out[f"gnorm/{k}"] = np.sqrt((g ** 2).sum())
        out[f"gsum/{k}"] = g.sum()
        if full_grads or g.size <= 4096:
            out[f"gfull/{k}"] = g
        else:
            idx = sample_idx(g.size)
            out[f"gidx/{k}"] = idx
            out[f"gval/{k}"] = g.reshape(-1)[idx]
    return out


CASES = {
    # name: (param_seed, batch_seed, batch, latent, hidden, layers, train, full_grads)
    "cfgb_full_b4": (101, 201, 4, 292, 501, 3, True, False),
    "cfgb_small_b3": (102, 202, 3, 16, 24, 3, True, True),
    "cfgb_small_eval_b2": (103, 203, 2, 8, 16, 2, False, True),
}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    for name, (ps, bs, B, Z, H, L, train, full) in CASES.items():
        P = vo.make_params(ps, dtype=np.float64, latent=Z, hidden=H, layers=L)
        ids, onehot, eps = vo.make_batch(bs, B, latent=Z, dtype=np.float64)
        r64 = run_reference(P, onehot, eps, Z, H, L, torch.float64, train)
        P32 = {k: v.astype(np.float32) for k, v in P.items()}
        r32 = run_reference(P32, onehot.astype(np.float32), eps.astype(np.float32), Z, H, L, torch.float32, train)
        out = {f"f64/{k}": v for k, v in pack(r64, full).items()}
        out.update({f"f32/{k}": v for k, v in pack(r32, full).items()})
        out["meta"] = np.array([ps, bs, B, Z, H, L, int(train)], dtype=np.int64)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "loss64", r64["loss"], "loss32", r32["loss"], "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
