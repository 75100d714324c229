"""Golden vectors for the MOSES VAE step from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden_moses.py

Imports /root/reference/mosesvae.py + vocab.py, loads seed-addressed parameters (oracle.moses_oracle.make_moses_params),
runs VAE.forward (mosesvae.py:126-140) in eval mode (dropout off) with eps injected through torch.randn_like, and
differentiates kl_weight*kl + recon (moses_train_distrib_logp.py:302-306)."""
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

CASES = {"moses_b6": (301, 401, 6, 0.1), "moses_b3_kl1": (302, 402, 3, 1.0)}


def main():
    import mosesvae
    import vocab as refvocab
    chars = [chr(ord("A") + i) for i in range(30)]
    voc = refvocab.OneHotVocab(chars)
    assert len(voc) == 34
    for name, (ps, bs, B, klw) in CASES.items():
        out = {}
        for tag, dt, tdt in (("f64", np.float64, torch.float64),):
            P = mo.make_moses_params(ps, dtype=np.float64)
            seqs, eps, pad = mo.make_moses_batch(bs, B, dtype=np.float64)
            assert pad == voc.pad
            model = mosesvae.VAE(voc).to(tdt)
            sd = model.state_dict()
            for k, v in P.items():
                sd[k].copy_(torch.from_numpy(v).to(tdt))      # aliases (encoder.N / decoder.N / vae.N) share storage
            model.eval()
            e = torch.from_numpy(eps).to(tdt)
            orig = torch.randn_like
            torch.randn_like = lambda t, *a, **k: e
            try:
                kl, recon, z, logvar, x, y = model([torch.from_numpy(s) for s in seqs])
            finally:
                torch.randn_like = orig
            loss = klw * kl + recon
            loss.backward()
            named = dict(model.named_parameters())
            out[f"{tag}/kl"], out[f"{tag}/recon"] = float(kl), float(recon)
            out[f"{tag}/z"] = z.detach().numpy().astype(np.float64)
            out[f"{tag}/y"] = y.detach().numpy().astype(np.float64)
            for k in P:
                g = named[k].grad.detach().numpy().astype(np.float64)
                out[f"{tag}/gnorm/{k}"] = np.sqrt((g ** 2).sum())
                if g.size <= 20000:
                    out[f"{tag}/gfull/{k}"] = g
                else:
                    rng = np.random.Generator(np.random.PCG64(7))
                    idx = np.sort(rng.choice(g.size, size=64, replace=False))
                    out[f"{tag}/gidx/{k}"], out[f"{tag}/gval/{k}"] = idx, g.reshape(-1)[idx]
        out["meta"] = np.array([ps, bs, B], dtype=np.int64)
        out["kl_weight"] = np.array([klw])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, out["f64/kl"], out["f64/recon"], os.path.getsize(path))


if __name__ == "__main__":
    main()
