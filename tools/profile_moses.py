"""Minimal driver for ncu: N direct (non-graph) fused MOSES VAE train steps (mosesvae.py shapes) at batch 4096, bf16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import molecular_vae_b200 as m
from oracle import moses_oracle as mo
from tests.test_gpu_moses import _Vocab

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = m.mosesvae.VAE(_Vocab(), precision="bf16").cuda()
seqs, eps, pad = mo.make_moses_batch(1, B)
x = [torch.from_numpy(s).cuda() for s in seqs]
eps = torch.from_numpy(eps).cuda()
for i in range(steps):
    out = model.elbo_step(x, kl_weight=0.1, eps=eps)
    torch.cuda.synchronize()
    print("step", i, out.cpu().numpy(), "launches", m._lib.lib.mvae_launch_count(), flush=True)
model.check_device_error()
