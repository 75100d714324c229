"""Context measurement (not part of the product, not used by bench.py): the same Config-B layer stack written with
stock torch.nn modules (cuDNN GRU / cuBLAS / ATen) -- i.e. what the reference's models2d.py + train.py:31-38 cost
under PyTorch eager on this GPU.  Prints molecules/s for fp32 (TF32 off/on) and bf16 autocast."""
import sys, time
import torch
from torch import nn
from torch.nn import functional as F


class RefVAE(nn.Module):  # layer stack of models2d.py:8-52 with latent 292
    def __init__(self, Z=292):
        super().__init__()
        self.conv1d1 = nn.Conv1d(120, 9, 9); self.conv1d2 = nn.Conv1d(9, 9, 9); self.conv1d3 = nn.Conv1d(9, 10, 11)
        self.fc0 = nn.Linear(90, 435); self.fc11 = nn.Linear(435, Z); self.fc12 = nn.Linear(435, Z)
        self.fc2 = nn.Linear(Z, Z); self.gru = nn.GRU(Z, 501, 3, batch_first=True); self.fc3 = nn.Linear(501, 35)

    def forward(self, x):
        h = F.relu(self.conv1d1(x)); h = F.relu(self.conv1d2(h)); h = F.relu(self.conv1d3(h))
        h = F.selu(self.fc0(h.view(h.size(0), -1)))
        mu, logvar = self.fc11(h), self.fc12(h)
        z = torch.randn_like(mu) * torch.exp(0.5 * logvar) + mu
        z = F.selu(self.fc2(z)).unsqueeze(1).repeat(1, 120, 1)
        out, _ = self.gru(z)
        y = F.softmax(self.fc3(out.contiguous().view(-1, out.size(-1))), dim=1).view(out.size(0), -1, 35)
        return y, mu, logvar


def loss_function(recon_x, x, mu, logvar, max_len=120):
    bce = max_len * F.binary_cross_entropy(recon_x.reshape(-1).float(), x.reshape(-1))
    return bce + (-0.5 * torch.mean(1. + mu - logvar ** 2. - torch.exp(mu)))


def run(B, mode, steps=5):
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    m = RefVAE().cuda()
    ids = torch.randint(0, 35, (B, 120), device="cuda")
    x = F.one_hot(ids, 35).float()
    def step():
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            y, mu, lv = m(x)
        loss = loss_function(y, x, mu.float(), lv.float())
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"torch eager {mode:5s} B={B}: {ms:8.2f} ms/step  {B / ms * 1e3:10.0f} molecules/s", flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for mode in ("fp32", "tf32", "bf16"):
        run(B, mode)
