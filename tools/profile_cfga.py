"""Minimal driver for ncu: N direct (non-graph) fused ELBO steps of Config A (models.py as shipped) at batch 4096, bf16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import molecular_vae_b200 as m
from oracle import vae_oracle as vo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(42)
model = m.models.MolecularVAE(precision="bf16").cuda()
ids, _, eps = vo.make_batch(1000, B)
ids, eps = torch.from_numpy(ids).cuda(), torch.from_numpy(eps).cuda()
for i in range(steps):
    out = model.elbo_step(ids, eps, max_len=120, use_graph=False)
    torch.cuda.synchronize()
    print("step", i, out.cpu().numpy(), "launches", m._lib.lib.mvae_launch_count())
model.engine(B).check_device_error()
