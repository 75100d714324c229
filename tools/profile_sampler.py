"""Minimal driver for ncu: one direct (non-graph) VAE.sample decode of 8192 latents (hugesample.py:113 batch), bf16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import molecular_vae_b200 as m
from tests.test_gpu_moses import _Vocab

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
model = m.mosesvae.VAE(_Vocab(), precision="bf16").cuda().eval()
z = torch.randn(B, 160, device="cuda")
for i in range(2):
    ids, lens, _ = model.sample_ids(B, max_len=L, z=z, greedy=True, use_graph=False)
    torch.cuda.synchronize()
    print("pass", i, "mean_len", lens.float().mean().item(), flush=True)
model.check_device_error()
