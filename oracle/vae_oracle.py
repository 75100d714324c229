"""CPU oracle for the molecular-VAE ELBO hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's forward + backward ELBO step
for the canonical conv / latent-292 / 3x501-GRU model ("Config B").  It is the
checker the CUDA path is compared against.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product
package never does (and has no CPU fallback).

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference modules themselves, run
in the build container on torch CPU (tests/golden/make_golden.py imports
/root/reference/models2d.py and train.py:31-38 and writes tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the oracle against those files).

Reference lines restated here:
  models2d.py:23-29   encode   : conv1d x3 (ReLU) -> flatten -> fc0 (SELU) -> fc11 / fc12
  models2d.py:31-38   reparametrize : z = mu + eps * exp(0.5 logvar) (train) | mu (eval)
  models2d.py:40-47   decode   : fc2 (SELU) -> repeat T -> GRU -> fc3 -> softmax over charset
  train.py:31-38      loss_function : max_len * BCE(mean) + swapped KL (mean)
  torch.nn.GRU        gate order r,z,n ; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))
  torch BCELoss       log terms clamped at >= -100 ; backward divides by max(p(1-p), 1e-12)
"""
from __future__ import annotations

import numpy as np

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


# ----------------------------------------------------------------------------
# parameter container helpers
# ----------------------------------------------------------------------------
def config_b_shapes(latent=292, hidden=501, layers=3, seq_len=120, charset=35):
    """state_dict key -> shape for the Config-B stack (models2d.py:12-21 with the
    latent widened from 2 to `latent`)."""
    l1 = charset - 9 + 1
    l2 = l1 - 9 + 1
    l3 = l2 - 11 + 1
    flat = 10 * l3
    s = {
        "conv1d1.weight": (9, seq_len, 9), "conv1d1.bias": (9,),
        "conv1d2.weight": (9, 9, 9), "conv1d2.bias": (9,),
        "conv1d3.weight": (10, 9, 11), "conv1d3.bias": (10,),
        "fc0.weight": (435, flat), "fc0.bias": (435,),
        "fc11.weight": (latent, 435), "fc11.bias": (latent,),
        "fc12.weight": (latent, 435), "fc12.bias": (latent,),
        "fc2.weight": (latent, latent), "fc2.bias": (latent,),
    }
    for l in range(layers):
        inp = latent if l == 0 else hidden
        s[f"gru.weight_ih_l{l}"] = (3 * hidden, inp)
        s[f"gru.weight_hh_l{l}"] = (3 * hidden, hidden)
        s[f"gru.bias_ih_l{l}"] = (3 * hidden,)
        s[f"gru.bias_hh_l{l}"] = (3 * hidden,)
    s["fc3.weight"] = (charset, hidden)
    s["fc3.bias"] = (charset,)
    return s


def make_params(seed, dtype=np.float32, scale=None, **cfg):
    """Deterministic synthetic parameters: U(-k, k) per tensor with k = 1/sqrt(fan_in)
    (the torch default-init law), drawn from numpy PCG64 so that tests, the golden
    generator and the GPU box all build identical weights without shipping them."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    shapes = config_b_shapes(**cfg)
    hidden = cfg.get("hidden", 501)
    for k, shp in shapes.items():
        if k.startswith("gru."):
            fan = hidden
        elif k.endswith(".weight"):
            fan = int(np.prod(shp[1:]))
        else:
            fan = int(np.prod(shapes[k.replace(".bias", ".weight")][1:]))
        bound = 1.0 / np.sqrt(fan) if scale is None else scale
        out[k] = rng.uniform(-bound, bound, size=shp).astype(dtype)
    return out


def make_batch(seed, batch, seq_len=120, charset=35, latent=292, dtype=np.float32):
    """ZINC-like synthetic batch: length ~ clip(round(N(44,9)),10,110)
    capped at seq_len, tokens uniform in 1..charset-1, pad id 0 (the space char,
    data_loader.py:27). Returns ids u8 (B,T), one-hot (B,T,C), eps (B,Z)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.clip(np.rint(rng.normal(44.0, 9.0, size=batch)), 10, min(110, seq_len)).astype(np.int64)
    ids = rng.integers(1, charset, size=(batch, seq_len), dtype=np.int64)
    pos = np.arange(seq_len)[None, :]
    ids = np.where(pos < lens[:, None], ids, 0).astype(np.uint8)
    onehot = np.zeros((batch, seq_len, charset), dtype=dtype)
    onehot[np.arange(batch)[:, None], pos, ids] = 1
    eps = rng.standard_normal(size=(batch, latent)).astype(dtype)
    return ids, onehot, eps


# ----------------------------------------------------------------------------
# elementwise pieces
# ----------------------------------------------------------------------------
def selu(x):
    # F.selu == scale * elu(x, alpha)   (models2d.py:28,41 ; models.py:58-68)
    return SELU_SCALE * np.where(x > 0, x, SELU_ALPHA * np.expm1(np.minimum(x, 0)))


def selu_grad(x):
    return SELU_SCALE * np.where(x > 0, 1.0, SELU_ALPHA * np.exp(np.minimum(x, 0))).astype(x.dtype)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def conv1d_valid(x, w, b):
    """nn.Conv1d, stride 1, no padding. x (B,Cin,L), w (Cout,Cin,K) -> (B,Cout,L-K+1)."""
    B, Cin, L = x.shape
    Cout, _, K = w.shape
    Lo = L - K + 1
    # im2col: (B, Lo, Cin*K)
    idx = np.arange(Lo)[:, None] + np.arange(K)[None, :]
    cols = x[:, :, idx]                       # (B,Cin,Lo,K)
    cols = cols.transpose(0, 2, 1, 3).reshape(B * Lo, Cin * K)
    y = cols @ w.reshape(Cout, Cin * K).T + b
    return y.reshape(B, Lo, Cout).transpose(0, 2, 1), cols


def conv1d_valid_bwd(dy, cols, w, x_shape, need_dx=True):
    B, Cin, L = x_shape
    Cout, _, K = w.shape
    Lo = L - K + 1
    dy2 = dy.transpose(0, 2, 1).reshape(B * Lo, Cout)
    dw = (dy2.T @ cols).reshape(w.shape)
    db = dy2.sum(0)
    dx = None
    if need_dx:
        dcols = (dy2 @ w.reshape(Cout, Cin * K)).reshape(B, Lo, Cin, K)
        dx = np.zeros(x_shape, dtype=dy.dtype)
        for k in range(K):
            dx[:, :, k:k + Lo] += dcols[:, :, :, k].transpose(0, 2, 1)
    return dx, dw, db


# ----------------------------------------------------------------------------
# GRU stack (torch.nn.GRU semantics, batch_first, h0 = 0)
# ----------------------------------------------------------------------------
def gru_stack_forward(x0, T, Ws, h0=None):
    """x0: (B,I) time-invariant input of layer 0 (the Repeat(T) of models2d.py:42 is
    never materialised: the layer-0 projection is computed once), or (B,T,I).
    Ws: list of (w_ih, w_hh, b_ih, b_hh). Returns top output (B,T,H) and a cache."""
    cache = []
    inp = x0
    for (w_ih, w_hh, b_ih, b_hh) in Ws:
        H = w_hh.shape[1]
        B = inp.shape[0]
        dt = w_hh.dtype
        if inp.ndim == 2:
            gi_all = (inp @ w_ih.T + b_ih)[:, None, :]          # broadcast over T
            gi_all = np.broadcast_to(gi_all, (B, T, 3 * H))
        else:
            gi_all = inp @ w_ih.T + b_ih
        h = np.zeros((B, H), dtype=dt) if h0 is None else h0
        hs = np.empty((B, T + 1, H), dtype=dt)
        hs[:, 0] = h
        r_s = np.empty((B, T, H), dtype=dt)
        z_s = np.empty((B, T, H), dtype=dt)
        n_s = np.empty((B, T, H), dtype=dt)
        ghn_s = np.empty((B, T, H), dtype=dt)
        for t in range(T):
            gi = gi_all[:, t]
            gh = h @ w_hh.T + b_hh
            r = sigmoid(gi[:, :H] + gh[:, :H])
            z = sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1.0 - z) * n + z * h
            hs[:, t + 1] = h
            r_s[:, t], z_s[:, t], n_s[:, t], ghn_s[:, t] = r, z, n, gh[:, 2 * H:]
        cache.append(dict(inp=inp, hs=hs, r=r_s, z=z_s, n=n_s, ghn=ghn_s))
        inp = hs[:, 1:]
    return inp, cache


def gru_stack_backward(dout, Ws, cache):
    """dout: (B,T,H) grad wrt top-layer outputs. Returns (dx0, [per-layer grads])."""
    grads = [None] * len(Ws)
    for l in reversed(range(len(Ws))):
        w_ih, w_hh, b_ih, b_hh = Ws[l]
        c = cache[l]
        B, T, H = c["r"].shape
        dt = w_hh.dtype
        dgi = np.empty((B, T, 3 * H), dtype=dt)
        dgh = np.empty((B, T, 3 * H), dtype=dt)
        dh = np.zeros((B, H), dtype=dt)
        for t in reversed(range(T)):
            dh = dh + dout[:, t]
            r, z, n, ghn = c["r"][:, t], c["z"][:, t], c["n"][:, t], c["ghn"][:, t]
            hprev = c["hs"][:, t]
            dn = dh * (1.0 - z)
            dz = dh * (hprev - n)
            dan = dn * (1.0 - n * n)
            dar = dan * ghn * r * (1.0 - r)
            daz = dz * z * (1.0 - z)
            dgi[:, t, :H], dgi[:, t, H:2 * H], dgi[:, t, 2 * H:] = dar, daz, dan
            dgh[:, t, :H], dgh[:, t, H:2 * H], dgh[:, t, 2 * H:] = dar, daz, dan * r
            dh = dh * z + dgh[:, t] @ w_hh
        hprev_all = c["hs"][:, :-1].reshape(B * T, H)
        dgh2 = dgh.reshape(B * T, 3 * H)
        dgi2 = dgi.reshape(B * T, 3 * H)
        g = {"w_hh": dgh2.T @ hprev_all, "b_hh": dgh2.sum(0), "b_ih": dgi2.sum(0)}
        inp = c["inp"]
        if inp.ndim == 2:
            dgi_sum = dgi.sum(1)                               # (B,3H)
            g["w_ih"] = dgi_sum.T @ inp
            dout = dgi_sum @ w_ih                              # (B,I)
        else:
            g["w_ih"] = dgi2.T @ inp.reshape(B * T, -1)
            dout = (dgi2 @ w_ih).reshape(B, T, -1)
        grads[l] = g
    return dout, grads


# ----------------------------------------------------------------------------
# loss pieces (train.py:31-38)
# ----------------------------------------------------------------------------
def softmax_rows(a):
    m = a.max(-1, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(-1, keepdims=True)


def bce_mean_times(p, x, max_len):
    """max_len * nn.BCELoss(mean)(p, x) with torch's -100 clamp on each log."""
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p), -100.0)
        l1p = np.maximum(np.log1p(-p), -100.0)
    return max_len * np.mean(-(x * lp + (1.0 - x) * l1p), dtype=np.float64)


def bce_grad_wrt_logits(p, x, max_len):
    """d(max_len*BCE_mean)/d(logits) through the softmax."""
    N = p.size
    g = (max_len / N) * (p - x) / np.maximum(p * (1.0 - p), 1e-12)
    return p * (g - (g * p).sum(-1, keepdims=True))


def kl_swapped(mu, logvar):
    """train.py:36-37 as shipped (mu and logvar swapped): -0.5*mean(1+mu-logvar^2-exp(mu))."""
    return -0.5 * np.mean(1.0 + mu - logvar ** 2 - np.exp(mu), dtype=np.float64)


# ----------------------------------------------------------------------------
# the whole step
# ----------------------------------------------------------------------------
def config_b_step(P, x_onehot, eps, max_len=120, train=True, need_grads=True, layers=3):
    """One full fwd(+bwd) ELBO step of Config B.

    P: dict state_dict-key -> ndarray.  x_onehot: (B,T,C) float.  eps: (B,Z).
    Returns dict with probs, mu, logvar, z, loss, bce, kl and (if need_grads) grads
    keyed like P."""
    dt = P["fc0.weight"].dtype
    x = x_onehot.astype(dt)
    B, T, C = x.shape
    # ---- encode (models2d.py:23-29); NB channels = sequence positions
    a1, cols1 = conv1d_valid(x, P["conv1d1.weight"], P["conv1d1.bias"])
    h1 = np.maximum(a1, 0)
    a2, cols2 = conv1d_valid(h1, P["conv1d2.weight"], P["conv1d2.bias"])
    h2 = np.maximum(a2, 0)
    a3, cols3 = conv1d_valid(h2, P["conv1d3.weight"], P["conv1d3.bias"])
    h3 = np.maximum(a3, 0)
    flat = h3.reshape(B, -1)
    a4 = flat @ P["fc0.weight"].T + P["fc0.bias"]
    h4 = selu(a4)
    mu = h4 @ P["fc11.weight"].T + P["fc11.bias"]
    logvar = h4 @ P["fc12.weight"].T + P["fc12.bias"]
    # ---- reparametrize (models2d.py:31-38)
    std = np.exp(0.5 * logvar)
    z = mu + eps.astype(dt) * std if train else mu
    # ---- decode (models2d.py:40-47)
    a5 = z @ P["fc2.weight"].T + P["fc2.bias"]
    zr = selu(a5)
    Ws = [(P[f"gru.weight_ih_l{l}"], P[f"gru.weight_hh_l{l}"],
           P[f"gru.bias_ih_l{l}"], P[f"gru.bias_hh_l{l}"]) for l in range(layers)]
    out, cache = gru_stack_forward(zr, T, Ws)
    logits = out @ P["fc3.weight"].T + P["fc3.bias"]          # (B,T,C)
    probs = softmax_rows(logits)                               # nn.Softmax() implicit dim=1 on (B*T,C)
    bce = bce_mean_times(probs, x, max_len)
    kl = kl_swapped(mu, logvar)
    res = dict(probs=probs, mu=mu, logvar=logvar, z=z, bce=float(bce), kl=float(kl),
               loss=float(bce + kl), argmax=probs.argmax(-1))
    if not need_grads:
        return res
    # ---- backward
    G = {}
    dlogits = bce_grad_wrt_logits(probs, x, max_len).astype(dt)
    dl2 = dlogits.reshape(B * T, C)
    G["fc3.weight"] = dl2.T @ out.reshape(B * T, -1)
    G["fc3.bias"] = dl2.sum(0)
    dout = (dl2 @ P["fc3.weight"]).reshape(B, T, -1)
    dzr, gg = gru_stack_backward(dout, Ws, cache)
    for l in range(layers):
        G[f"gru.weight_ih_l{l}"] = gg[l]["w_ih"]
        G[f"gru.weight_hh_l{l}"] = gg[l]["w_hh"]
        G[f"gru.bias_ih_l{l}"] = gg[l]["b_ih"]
        G[f"gru.bias_hh_l{l}"] = gg[l]["b_hh"]
    da5 = dzr * selu_grad(a5)
    G["fc2.weight"] = da5.T @ z
    G["fc2.bias"] = da5.sum(0)
    dz = da5 @ P["fc2.weight"]
    BZ = mu.size
    dmu = dz + (-0.5 * (1.0 - np.exp(mu)) / BZ)
    dlv = logvar / BZ
    if train:
        dlv = dlv + dz * eps.astype(dt) * std * 0.5
    dmu = dmu.astype(dt)
    dlv = dlv.astype(dt)
    G["fc11.weight"] = dmu.T @ h4
    G["fc11.bias"] = dmu.sum(0)
    G["fc12.weight"] = dlv.T @ h4
    G["fc12.bias"] = dlv.sum(0)
    dh4 = dmu @ P["fc11.weight"] + dlv @ P["fc12.weight"]
    da4 = dh4 * selu_grad(a4)
    G["fc0.weight"] = da4.T @ flat
    G["fc0.bias"] = da4.sum(0)
    dh3 = (da4 @ P["fc0.weight"]).reshape(h3.shape)
    da3 = dh3 * (a3 > 0)
    dh2, G["conv1d3.weight"], G["conv1d3.bias"] = conv1d_valid_bwd(da3, cols3, P["conv1d3.weight"], h2.shape)
    da2 = dh2 * (a2 > 0)
    dh1, G["conv1d2.weight"], G["conv1d2.bias"] = conv1d_valid_bwd(da2, cols2, P["conv1d2.weight"], h1.shape)
    da1 = dh1 * (a1 > 0)
    _, G["conv1d1.weight"], G["conv1d1.bias"] = conv1d_valid_bwd(da1, cols1, P["conv1d1.weight"], x.shape, need_dx=False)
    res["grads"] = G
    return res


def greedy_decode(P, z, T=120, layers=3):
    """Greedy decode of fixed latents (train.py:110 / train_sample.py:31-37 applied to
    the Config-B decoder): argmax over the charset at every position."""
    a5 = z @ P["fc2.weight"].T + P["fc2.bias"]
    zr = selu(a5)
    Ws = [(P[f"gru.weight_ih_l{l}"], P[f"gru.weight_hh_l{l}"],
           P[f"gru.bias_ih_l{l}"], P[f"gru.bias_hh_l{l}"]) for l in range(layers)]
    out, _ = gru_stack_forward(zr, T, Ws)
    logits = out @ P["fc3.weight"].T + P["fc3.bias"]
    return logits.argmax(-1).astype(np.uint8), logits
