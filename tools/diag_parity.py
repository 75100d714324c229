"""Prints the measured parity errors (loss terms, worst gradient tensors) of the CUDA paths against the oracle / fixtures:
the numbers the tolerances in tests/ are set from.  python tools/diag_parity.py [moses] [cfgb]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import molecular_vae_b200 as m  # noqa: E402
from oracle import moses_oracle as mo  # noqa: E402
from tests.test_gpu_moses import _setup, _setup_file  # noqa: E402
from tests.util_gpu import build_model, grads_of, make_case, rel_l2  # noqa: E402


def moses():
    for prec, B in (("fp32", 6), ("fp32", 70), ("bf16", 64), ("bf16", 300)):
        P, seqs, eps, pad, model = _setup(m, prec, 311, 411 + B, B)
        ref = mo.moses_step({k: v.astype(np.float64) for k, v in P.items()}, seqs, eps.astype(np.float64), pad, kl_weight=0.1)
        out = model.elbo_step([torch.from_numpy(s).cuda() for s in seqs], kl_weight=0.1, eps=torch.from_numpy(eps).cuda())
        sc = out.cpu().numpy()
        errs = {k: rel_l2(p.grad.cpu().numpy(), ref["grads"][k]) for k, p in model.named_parameters() if k in ref["grads"]}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"moses {prec} B={B}: kl {abs(sc[1]-ref['kl'])/abs(ref['kl']):.2e} recon {abs(sc[2]-ref['recon'])/abs(ref['recon']):.2e} worst", worst, flush=True)
    for prec, B in (("fp32", 7), ("fp32", 70), ("bf16", 200)):
        P, seqs, eps, pad, model = _setup_file(m, prec, 331, 431 + B, B)
        ref = mo.mosesfile_step({k: v.astype(np.float64) for k, v in P.items()}, seqs, eps.astype(np.float64), pad, kl_weight=0.5)
        out = model.elbo_step([torch.from_numpy(s).cuda() for s in seqs], kl_weight=0.5, eps=torch.from_numpy(eps).cuda())
        sc = out.cpu().numpy()
        errs = {k: rel_l2(p.grad.cpu().numpy(), ref["grads"][k]) for k, p in model.named_parameters() if k in ref["grads"]}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"mosesfile {prec} B={B}: kl {abs(sc[1]-ref['kl'])/abs(ref['kl']):.2e} recon {abs(sc[2]-ref['recon'])/abs(ref['recon']):.2e} worst", worst, flush=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "moses_b4096.npz"))
    ps, bs, B, M = [int(v) for v in g["meta"]]
    for train in (False,):
        P, seqs, eps, pad, model = _setup(m, "bf16", ps, bs, B)
        out = model.elbo_step([torch.from_numpy(s).cuda() for s in seqs], kl_weight=float(g["kl_weight"][0]), eps=torch.from_numpy(eps).cuda())
        sc = out.cpu().numpy()
        errs = {}
        for k, p in model.named_parameters():
            if f"f64/gnorm/{k}" not in g:
                continue
            gr = p.grad.cpu().numpy()
            errs[k] = rel_l2(gr, g[f"f64/gfull/{k}"]) if f"f64/gfull/{k}" in g else rel_l2(gr.reshape(-1)[g[f"f64/gidx/{k}"]], g[f"f64/gval/{k}"])
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print(f"moses bf16 B=4096 vs reference fixture: kl {abs(sc[1]-g['f64/kl'])/abs(g['f64/kl']):.2e} recon {abs(sc[2]-g['f64/recon'])/abs(g['f64/recon']):.2e} M {sc[3]} {M} worst", worst, flush=True)


def cfgb():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfgb_full_b4096.npz"))
    ps, bs, B, Z, H, L, train = [int(v) for v in g["meta"]]
    P, ids, onehot, eps = make_case(ps, bs, B, Z, H, L)
    model = build_model(m, P, Z, H, L, "bf16")
    out = model.elbo_step(torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda(), max_len=120, use_graph=False)
    sc = out.cpu().numpy()
    errs = {}
    for k, gr in grads_of(model).items():
        errs[k] = rel_l2(gr, g[f"f64/gfull/{k}"]) if f"f64/gfull/{k}" in g else rel_l2(gr.reshape(-1)[g[f"f64/gidx/{k}"]], g[f"f64/gval/{k}"])
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"cfgb bf16 B=4096 vs reference fixture: loss {abs(sc[0]-g['f64/loss'])/abs(g['f64/loss']):.2e} bce {abs(sc[1]-g['f64/bce'])/abs(g['f64/bce']):.2e} "
          f"kl {abs(sc[2]-g['f64/kl'])/abs(g['f64/kl']):.2e} worst", worst, flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["moses", "cfgb"]
    if "moses" in which:
        moses()
    if "cfgb" in which:
        cfgb()
