"""Loss-curve fixtures from the UNMODIFIED reference modules (build container only).

    python tests/golden/make_golden_curves.py

Runs the loops of tests/curves.py (train.py:94-104 and moses_train_distrib_logp.py:289-338 restated) over
/root/reference/models2d.py (latent 292) + train.py:31-38 and /root/reference/mosesvae.py on torch CPU, in float64 (the
anchor) and float32 (what the reference runs in: the spread between the two is the budget a 1e-4 per-step bound is read
against), and stores the per-step losses."""
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
from tests import curves  # noqa: E402
import make_golden as mg  # noqa: E402


class EpsPatch:
    """torch.randn_like -> the injected draws, for the duration of the loop (the reference modules are not edited)."""

    def __init__(self, tdt):
        self.tdt, self.eps, self.orig = tdt, None, torch.randn_like

    def __enter__(self):
        torch.randn_like = lambda t, *a, **k: self.eps
        return self

    def __exit__(self, *a):
        torch.randn_like = self.orig

    def set(self, eps):
        self.eps = torch.from_numpy(np.asarray(eps)).to(self.tdt)


def cfgb_curve(tdt):
    c = curves.CFGB
    ndt = np.float64 if tdt == torch.float64 else np.float32
    P = {k: v.astype(ndt) for k, v in vo.make_params(c["param_seed"], dtype=np.float64, latent=c["Z"], hidden=c["H"], layers=c["L"]).items()}
    model = mg.build_reference_model(P, c["Z"], c["H"], c["L"], tdt).train()
    lf = mg.load_loss_function(c["max_len"])
    with EpsPatch(tdt) as ep:
        return curves.cfgb_loop(model, lf, lambda onehot: torch.from_numpy(onehot).to(tdt), ep.set)


def moses_curve(tdt):
    import mosesvae
    import vocab as refvocab
    c = curves.MOSES
    voc = refvocab.OneHotVocab([chr(ord("A") + i) for i in range(30)])
    P = mo.make_moses_params(c["param_seed"], dtype=np.float64)
    model = mosesvae.VAE(voc).to(tdt)
    sd = model.state_dict()
    for k, v in P.items():
        sd[k].copy_(torch.from_numpy(v).to(tdt))
    model.eval()
    with EpsPatch(tdt) as ep:
        return curves.moses_loop(model, lambda t: t, ep.set)


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    for tag, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        out[f"cfgb/{tag}"] = cfgb_curve(tdt)
        print("cfgb", tag, out[f"cfgb/{tag}"][[0, 1, -1]], flush=True)
        a, mn, kl = moses_curve(tdt)
        out[f"moses_agg/{tag}"], out[f"moses_main/{tag}"], out[f"moses_kl/{tag}"] = a, mn, kl
        print("moses", tag, a[[0, -1]], mn[[0, -1]], kl[[0, -1]], flush=True)
    for k in ("cfgb", "moses_agg", "moses_main", "moses_kl"):
        print(k, "max rel spread fp32 vs fp64:", float(np.max(np.abs(out[k + "/f32"] - out[k + "/f64"]) / np.abs(out[k + "/f64"]))))
    np.savez_compressed(os.path.join(HERE, "loss_curves.npz"), **out)


if __name__ == "__main__":
    main()
