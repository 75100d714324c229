"""Golden vectors for the BindingModel property head from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_binding.py

Imports /root/reference/mosesvae.py, loads seed-addressed parameters / BatchNorm buffers (oracle.binding_oracle), runs
BindingModel.forward (mosesvae.py:24-25) in train and eval mode on seeded latents and back-propagates a seeded upstream
gradient; stores the output, dz, the updated running statistics and every parameter gradient (float64)."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
from oracle import binding_oracle as bo  # noqa: E402


def main():
    import mosesvae
    out = {}
    B, Z = 12, 128
    rng = np.random.Generator(np.random.PCG64(77))
    z = rng.standard_normal((B, Z))
    dout = rng.standard_normal(B)
    out["z"], out["dout"], out["meta"] = z, dout, np.array([501, B, Z])
    for mode in ("train", "eval"):
        P, run = bo.make_binding_params(501, Z, dtype=np.float64)
        m = mosesvae.BindingModel(Z).double()
        sd = m.state_dict()
        for k, v in P.items():
            sd["binding_model." + k].copy_(torch.from_numpy(v))
        for k, v in run.items():
            sd["binding_model." + k].copy_(torch.from_numpy(v))
        m.train(mode == "train")
        zt = torch.from_numpy(z).requires_grad_(True)
        o = m(zt)
        o.backward(torch.from_numpy(dout).view(B, 1))
        out[f"{mode}/out"] = o.detach().numpy()
        out[f"{mode}/dz"] = zt.grad.numpy()
        for k, p in m.binding_model.named_parameters():
            out[f"{mode}/grad/{k}"] = p.grad.numpy()
        for k in run:
            out[f"{mode}/{k}"] = m.state_dict()["binding_model." + k].numpy()
    path = os.path.join(HERE, "binding_b12.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
