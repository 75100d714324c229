"""CPU oracle for the BindingModel property head (TEST INFRASTRUCTURE ONLY).

numpy restatement (+ hand-derived backward) of mosesvae.py:6-25:
  Linear(Z,256) -> BatchNorm1d(256) -> Tanh -> Linear(256,256) -> ReLU -> Linear(256,64) -> BatchNorm1d(64) -> ReLU -> Linear(64,1)
BatchNorm1d as torch: train mode normalises with the batch mean / biased variance and moves the running estimates with
momentum 0.1 towards the batch mean / UNBIASED variance; eval mode uses the running estimates.
Pinned by tests/golden/make_golden_binding.py (imports /root/reference/mosesvae.py) -> tests/golden/binding_*.npz."""
import numpy as np

KEYS = ["0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "5.weight", "5.bias", "6.weight", "6.bias",
        "8.weight", "8.bias"]


def binding_shapes(Z=128):
    return {"0.weight": (256, Z), "0.bias": (256,), "1.weight": (256,), "1.bias": (256,), "3.weight": (256, 256),
            "3.bias": (256,), "5.weight": (64, 256), "5.bias": (64,), "6.weight": (64,), "6.bias": (64,), "8.weight": (1, 64),
            "8.bias": (1,)}


def make_binding_params(seed, Z=128, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    P = {}
    for k, shp in binding_shapes(Z).items():
        if k in ("1.weight", "6.weight"):
            P[k] = rng.uniform(0.5, 1.5, size=shp).astype(dtype)
        elif k in ("1.bias", "6.bias"):
            P[k] = rng.uniform(-0.3, 0.3, size=shp).astype(dtype)
        else:
            fan = shp[-1] if len(shp) > 1 else binding_shapes(Z)[k.replace("bias", "weight")][-1]
            b = 1.0 / np.sqrt(fan)
            P[k] = rng.uniform(-b, b, size=shp).astype(dtype)
    run = {"1.running_mean": rng.normal(0, 0.2, 256).astype(dtype), "1.running_var": rng.uniform(0.5, 1.5, 256).astype(dtype),
           "6.running_mean": rng.normal(0, 0.2, 64).astype(dtype), "6.running_var": rng.uniform(0.5, 1.5, 64).astype(dtype)}
    return P, run


def _bn_fwd(x, g, b, rm, rv, train, eps, mom):
    if train:
        mean, var = x.mean(0), x.var(0)
        n = x.shape[0]
        new_rm = (1 - mom) * rm + mom * mean
        new_rv = (1 - mom) * rv + mom * var * n / (n - 1)
    else:
        mean, var, new_rm, new_rv = rm, rv, rm, rv
    inv = 1.0 / np.sqrt(var + eps)
    xh = (x - mean) * inv
    return g * xh + b, xh, inv, new_rm, new_rv


def _bn_bwd(dv, xh, inv, g, train):
    dg, db = (dv * xh).sum(0), dv.sum(0)
    if train:
        dx = g * inv * (dv - dv.mean(0) - xh * (dv * xh).mean(0))
    else:
        dx = g * inv * dv
    return dx, dg, db


def binding_step(P, run, z, dout=None, train=True, eps=1e-5, mom=0.1):
    a1 = z @ P["0.weight"].T + P["0.bias"]
    y1, xh1, inv1, rm1, rv1 = _bn_fwd(a1, P["1.weight"], P["1.bias"], run["1.running_mean"], run["1.running_var"], train, eps, mom)
    t1 = np.tanh(y1)
    r2 = np.maximum(t1 @ P["3.weight"].T + P["3.bias"], 0)
    a3 = r2 @ P["5.weight"].T + P["5.bias"]
    y3, xh3, inv3, rm3, rv3 = _bn_fwd(a3, P["6.weight"], P["6.bias"], run["6.running_mean"], run["6.running_var"], train, eps, mom)
    r3 = np.maximum(y3, 0)
    out = r3 @ P["8.weight"].T + P["8.bias"]
    res = dict(out=out, running={"1.running_mean": rm1, "1.running_var": rv1, "6.running_mean": rm3, "6.running_var": rv3})
    if dout is None:
        return res
    G = {}
    do = dout.reshape(-1, 1)
    G["8.weight"] = do.T @ r3; G["8.bias"] = do.sum(0)
    d3 = (do @ P["8.weight"]) * (y3 > 0)
    d3, G["6.weight"], G["6.bias"] = _bn_bwd(d3, xh3, inv3, P["6.weight"], train)
    G["5.weight"] = d3.T @ r2; G["5.bias"] = d3.sum(0)
    d2 = (d3 @ P["5.weight"]) * (r2 > 0)
    G["3.weight"] = d2.T @ t1; G["3.bias"] = d2.sum(0)
    d1 = (d2 @ P["3.weight"]) * (1 - t1 * t1)
    d1, G["1.weight"], G["1.bias"] = _bn_bwd(d1, xh1, inv1, P["1.weight"], train)
    G["0.weight"] = d1.T @ z; G["0.bias"] = d1.sum(0)
    res["grads"] = G
    res["dz"] = d1 @ P["0.weight"]
    return res
