"""Runs the UNMODIFIED reference (baseline/_ref/, see make_ref.py) for bench.py: models2d.VAE widened to latent 292 the
way SURVEY.md 8c describes (attribute replacement after construction -- the module file is not edited) + loss_function
AST-extracted from train.py:31-38 (the script imports comet_ml and reads absolute paths at import time).  Nothing of this
repository's kernels, models or engine is on this path."""
import ast
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def have_ref():
    return os.path.exists(os.path.join(REF, "models2d.py")) and os.path.exists(os.path.join(REF, "train.py"))


def load_loss_function(max_len=120):
    import torch
    tree = ast.parse(open(os.path.join(REF, "train.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "loss_function"][0]
    ns = {"torch": torch, "nn": torch.nn, "max_len": max_len}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train.py[loss_function]", "exec"), ns)
    return ns["loss_function"]


def build_cfgb(latent=292, hidden=501, layers=3, seed=42):
    import torch
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import models2d
    torch.manual_seed(seed)
    m = models2d.VAE()
    nn = torch.nn
    m.fc11, m.fc12 = nn.Linear(435, latent), nn.Linear(435, latent)
    m.fc2 = nn.Linear(latent, latent)
    m.gru = nn.GRU(latent, hidden, layers, batch_first=True)
    m.fc3 = nn.Linear(hidden, 35)
    return m


def synthetic_onehot(batch, seed=43):
    """ZINC-like ids (length ~ clip(N(44,9),10,110), ids 1..34, pad id 0: SURVEY.md 8d) as the float one-hot the reference
    consumes (data_loader.py:26-31)."""
    import numpy as np
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.clip(np.rint(rng.normal(44.0, 9.0, size=batch)), 10, 110).astype(np.int64)
    ids = rng.integers(1, 35, size=(batch, 120))
    ids[np.arange(120)[None, :] >= lens[:, None]] = 0
    return torch.nn.functional.one_hot(torch.from_numpy(ids), 35).float()


def step_rate(batch, steps, warmup, device="cpu", autocast=False, threads=None):
    """molecules/s of reference forward + loss_function + backward (train.py:98-101) on `device`."""
    import torch
    if device == "cpu":
        # torchrun exports OMP_NUM_THREADS=1 to every rank: ask for all host cores explicitly so the CPU arm is the same at any N
        torch.set_num_threads(threads or os.cpu_count() or 1)
    model = build_cfgb().to(device).train()
    lf = load_loss_function(120)
    x = synthetic_onehot(batch).to(device)

    def one():
        model.zero_grad(set_to_none=True)
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                probs, mu, logvar = model(x)
            loss = lf(probs.float(), x, mu.float(), logvar.float())
        else:
            probs, mu, logvar = model(x)
            loss = lf(probs, x, mu, logvar)
        loss.backward()
        return loss

    for _ in range(warmup):
        one()
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = one()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, float(loss)
