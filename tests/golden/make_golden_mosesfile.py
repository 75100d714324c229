"""Golden vectors for the bidirectional MOSES VAE variant from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_mosesfile.py

Imports /root/reference/mosesfile.py + config.py + vocab.py: mosesfile.VAE(vocab, config) with `--q_bidir` (the only setting
in which the shipped class runs: its encoder is hard-coded bidirectional, mosesfile.py:21-28), d_z = 128 (config.py:34-36).
Seed-addressed parameters (oracle.moses_oracle.make_mosesfile_params); eps injected through torch.randn_like; the loss that
is differentiated is kl_weight*kl + recon.  forward returns only (kl, recon) (mosesfile.py:100), so z is captured from
forward_encoder."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from oracle import moses_oracle as mo  # noqa: E402

CASES = {"mosesfile_b5": (321, 421, 5, 0.5)}


def main():
    import config as refconfig
    import mosesfile
    import vocab as refvocab
    voc = refvocab.OneHotVocab([chr(ord("A") + i) for i in range(30)])
    cfg = refconfig.get_parser().parse_args(["--q_bidir"])
    for name, (ps, bs, B, klw) in CASES.items():
        P = mo.make_mosesfile_params(ps, dtype=np.float64)
        seqs, eps, pad = mo.make_moses_batch(bs, B, d_z=128, dtype=np.float64)
        model = mosesfile.VAE(voc, cfg).double()
        sd = model.state_dict()
        for k, v in P.items():
            sd[k].copy_(torch.from_numpy(v))
        model.eval()
        e = torch.from_numpy(eps)
        orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: e
        try:
            x = [torch.from_numpy(s) for s in seqs]
            z, _ = model.forward_encoder(x)
            kl, recon = model(x)
        finally:
            torch.randn_like = orig
        (klw * kl + recon).backward()
        named = dict(model.named_parameters())
        out = {"f64/kl": float(kl), "f64/recon": float(recon), "f64/z": z.detach().numpy(),
               "meta": np.array([ps, bs, B], dtype=np.int64), "kl_weight": np.array([klw])}
        for k in P:
            g = named[k].grad.detach().numpy()
            out[f"f64/gnorm/{k}"] = np.sqrt((g ** 2).sum())
            if g.size <= 20000:
                out[f"f64/gfull/{k}"] = g
            else:
                idx = np.sort(np.random.Generator(np.random.PCG64(7)).choice(g.size, size=64, replace=False))
                out[f"f64/gidx/{k}"], out[f"f64/gval/{k}"] = idx, g.reshape(-1)[idx]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, out["f64/kl"], out["f64/recon"], os.path.getsize(path))


if __name__ == "__main__":
    main()
