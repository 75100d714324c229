"""Golden vectors at the BENCHMARKED batch size (BASELINE.json configs[1]/[3]: batch 4096) from the UNMODIFIED reference
modules (run in the build container only: /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_large.py [cfgb] [moses]

The reference's losses are batch means, so the full-batch loss / gradient is the weighted sum of what the reference
computes on contiguous chunks of the batch (weights n_c/B for the per-molecule means, M_c/M for the MOSES CE whose
normaliser is the global count of non-pad targets, mosesvae.py:193-197).  Each chunk goes through the reference's own
forward + loss + autograd (models2d.py:8-52 + train.py:31-38, mosesvae.py:126-199) exactly as make_golden.py /
make_golden_moses.py do for the small cases; chunking only bounds the memory of torch CPU autograd at B=4096.
Stored per parameter tensor: L2 norm, sum, and the gradient itself (<= 4096 entries) or 4096 sampled entries, in
float64 (the anchor) and float32 (what the reference runs in -> the measured fp32-vs-fp64 error budget `f32err/*`).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from oracle import moses_oracle as mo  # noqa: E402
from oracle import vae_oracle as vo  # noqa: E402
import make_golden as mg  # noqa: E402

NSAMPLE = 4096


def sample_idx(n, seed=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.sort(rng.choice(n, size=min(NSAMPLE, n), replace=False))


def pack_grads(out, tag, grads):
    for k, g in grads.items():
        out[f"{tag}/gnorm/{k}"] = np.sqrt((g ** 2).sum())
        out[f"{tag}/gsum/{k}"] = g.sum()
        if g.size <= NSAMPLE:
            out[f"{tag}/gfull/{k}"] = g
        else:
            idx = sample_idx(g.size)
            out[f"{tag}/gidx/{k}"] = idx
            out[f"{tag}/gval/{k}"] = g.reshape(-1)[idx]


def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))


# ---------------------------------------------------------------------------------------------------------------
def cfgb_case(ps, bs, B, chunk, tdt):
    Z, H, L, max_len = 292, 501, 3, 120
    ndt = np.float64 if tdt == torch.float64 else np.float32
    P = {k: v.astype(ndt) for k, v in vo.make_params(ps, dtype=np.float64, latent=Z, hidden=H, layers=L).items()}
    ids, onehot, eps = vo.make_batch(bs, B, latent=Z, dtype=np.float64)
    m = mg.build_reference_model(P, Z, H, L, tdt)
    m.train(True)
    lf = mg.load_loss_function(max_len)
    tot = dict(loss=0.0, bce=0.0, kl=0.0)
    mus = []
    for c0 in range(0, B, chunk):
        x = torch.from_numpy(onehot[c0:c0 + chunk]).to(tdt)
        e = torch.from_numpy(eps[c0:c0 + chunk]).to(tdt)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: e
        try:
            probs, mu, logvar = m(x)
        finally:
            torch.randn_like = orig
        w = x.shape[0] / B
        loss = lf(probs, x, mu, logvar)
        bce = max_len * torch.nn.functional.binary_cross_entropy(probs.reshape(-1), x.reshape(-1))
        (loss * w).backward()
        tot["loss"] += float(loss) * w
        tot["bce"] += float(bce) * w
        tot["kl"] += float(loss - bce) * w
        mus.append(mu.detach().numpy().astype(np.float64))
        print("  cfgb", tdt, c0, flush=True)
    grads = {k: p.grad.detach().numpy().astype(np.float64) for k, p in m.named_parameters()}
    return tot, grads, np.concatenate(mus)


def make_cfgb(name="cfgb_full_b4096", ps=111, bs=211, B=4096, chunk=512):
    out = {}
    res = {}
    for tag, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        tot, grads, mu = cfgb_case(ps, bs, B, chunk, tdt)
        res[tag] = grads
        for k, v in tot.items():
            out[f"{tag}/{k}"] = v
        out[f"{tag}/mu_head"] = mu[:8]
        pack_grads(out, tag, grads)
    for k in res["f64"]:
        out[f"f32err/{k}"] = rel_l2(res["f32"][k], res["f64"][k])
    out["meta"] = np.array([ps, bs, B, 292, 501, 3, 1], dtype=np.int64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, out["f64/loss"], out["f32/loss"], "max fp32-vs-fp64 grad err",
          max(out[f"f32err/{k}"] for k in res["f64"]), os.path.getsize(path))


# ---------------------------------------------------------------------------------------------------------------
def moses_case(ps, bs, B, chunk, klw, tdt):
    import mosesvae
    import vocab as refvocab
    voc = refvocab.OneHotVocab([chr(ord("A") + i) for i in range(30)])
    P = mo.make_moses_params(ps, dtype=np.float64)
    seqs, eps, pad = mo.make_moses_batch(bs, B, dtype=np.float64)
    assert pad == voc.pad
    model = mosesvae.VAE(voc).to(tdt)
    sd = model.state_dict()
    for k, v in P.items():
        sd[k].copy_(torch.from_numpy(v).to(tdt))
    model.eval()
    M = sum(len(s) - 1 for s in seqs)
    kl_tot = recon_tot = 0.0
    for c0 in range(0, B, chunk):
        part = seqs[c0:c0 + chunk]
        e = torch.from_numpy(eps[c0:c0 + chunk]).to(tdt)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: e
        try:
            kl, recon, z, logvar, x, y = model([torch.from_numpy(s) for s in part])
        finally:
            torch.randn_like = orig
        wk = len(part) / B
        wr = sum(len(s) - 1 for s in part) / M
        (klw * wk * kl + wr * recon).backward()
        kl_tot += float(kl) * wk
        recon_tot += float(recon) * wr
        print("  moses", tdt, c0, flush=True)
    named = dict(model.named_parameters())
    grads = {k: named[k].grad.detach().numpy().astype(np.float64) for k in P}
    return kl_tot, recon_tot, M, grads


def make_moses(name="moses_b4096", ps=341, bs=441, B=4096, chunk=512, klw=0.1):
    out = {}
    res = {}
    for tag, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        kl, recon, M, grads = moses_case(ps, bs, B, chunk, klw, tdt)
        res[tag] = grads
        out[f"{tag}/kl"], out[f"{tag}/recon"] = kl, recon
        pack_grads(out, tag, grads)
    for k in res["f64"]:
        out[f"f32err/{k}"] = rel_l2(res["f32"][k], res["f64"][k])
    out["meta"] = np.array([ps, bs, B, M], dtype=np.int64)
    out["kl_weight"] = np.array([klw])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, out["f64/kl"], out["f64/recon"], "max fp32-vs-fp64 grad err",
          max(out[f"f32err/{k}"] for k in res["f64"]), os.path.getsize(path))


def moses_budget():
    """fp32-vs-fp64 deviation of the REFERENCE itself on the small MOSES fixtures (the error budget the fp32 check mode is
    held to): re-runs make_golden_moses' cases in float32 and stores rel-L2 per tensor."""
    import mosesvae
    import vocab as refvocab
    voc = refvocab.OneHotVocab([chr(ord("A") + i) for i in range(30)])
    out = {}
    for name, (ps, bs, B, klw) in {"moses_b6": (301, 401, 6, 0.1), "moses_b70": (311, 481, 70, 0.1)}.items():
        g = {}
        for tag, tdt in (("f64", torch.float64), ("f32", torch.float32)):
            P = mo.make_moses_params(ps, dtype=np.float64)
            seqs, eps, pad = mo.make_moses_batch(bs, B, dtype=np.float64)
            model = mosesvae.VAE(voc).to(tdt)
            sd = model.state_dict()
            for k, v in P.items():
                sd[k].copy_(torch.from_numpy(v).to(tdt))
            model.eval()
            e = torch.from_numpy(eps).to(tdt)
            orig = torch.randn_like
            torch.randn_like = lambda t, *a, **k: e
            try:
                kl, recon, *_ = model([torch.from_numpy(s) for s in seqs])
            finally:
                torch.randn_like = orig
            (klw * kl + recon).backward()
            named = dict(model.named_parameters())
            g[tag] = ({k: named[k].grad.detach().numpy().astype(np.float64) for k in P}, float(kl), float(recon))
        for k in g["f64"][0]:
            out[f"{name}/f32err/{k}"] = rel_l2(g["f32"][0][k], g["f64"][0][k])
        out[f"{name}/f32err/kl"] = abs(g["f32"][1] - g["f64"][1]) / abs(g["f64"][1])
        out[f"{name}/f32err/recon"] = abs(g["f32"][2] - g["f64"][2]) / abs(g["f64"][2])
        print(name, "reference fp32 vs fp64: kl", out[f"{name}/f32err/kl"], "recon", out[f"{name}/f32err/recon"],
              "max grad", max(v for k, v in out.items() if k.startswith(name + "/f32err/") and "." in k.split("/")[-1]))
    np.savez_compressed(os.path.join(HERE, "moses_fp32_budget.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["cfgb", "moses", "budget"]
    if "budget" in which:
        moses_budget()
    if "cfgb" in which:
        make_cfgb()
    if "moses" in which:
        make_moses()
