"""The oracle (oracle/vae_oracle.py) against the fixtures produced by the reference
modules themselves (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import vae_oracle as vo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["cfgb_full_b4", "cfgb_small_b3", "cfgb_small_eval_b2"]


def _run(name, dtype):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ps, bs, B, Z, H, L, train = [int(v) for v in g["meta"]]
    P = vo.make_params(ps, dtype=np.float64, latent=Z, hidden=H, layers=L)
    ids, onehot, eps = vo.make_batch(bs, B, latent=Z, dtype=np.float64)
    P = {k: v.astype(dtype) for k, v in P.items()}
    r = vo.config_b_step(P, onehot.astype(dtype), eps.astype(dtype), max_len=120, train=bool(train), layers=L)
    return g, r


def _check(g, r, pre, rtol_loss, rtol_grad):
    assert abs(r["loss"] - g[pre + "loss"]) <= rtol_loss * abs(g[pre + "loss"])
    assert abs(r["bce"] - g[pre + "bce"]) <= rtol_loss * abs(g[pre + "bce"])
    # the f32 fixture's kl is (loss - bce) evaluated in fp32: budget it against |loss|
    assert abs(r["kl"] - g[pre + "kl"]) <= rtol_loss * abs(g[pre + "kl"]) + (1e-12 if pre == "f64/" else 2e-7 * abs(g[pre + "loss"]))
    np.testing.assert_allclose(r["mu"], g[pre + "mu"], rtol=rtol_grad, atol=rtol_grad)
    np.testing.assert_allclose(r["probs"], g[pre + "probs"], rtol=rtol_grad, atol=rtol_grad)
    for k, gr in r["grads"].items():
        gn = float(g[f"{pre}gnorm/{k}"])
        assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - gn) <= rtol_grad * gn + 1e-12, k
        if f"{pre}gfull/{k}" in g:
            ref = g[f"{pre}gfull/{k}"]
            err = np.abs(gr - ref).max()
        else:
            ref = g[f"{pre}gval/{k}"]
            err = np.abs(gr.reshape(-1)[g[f"{pre}gidx/{k}"]] - ref).max()
        # error relative to the tensor's rms-scale
        scale = gn / np.sqrt(gr.size) + 1e-30
        assert err <= rtol_grad * max(scale, np.abs(ref).max()), (k, err, scale)


@pytest.mark.parametrize("name", CASES)
def test_oracle_f64_matches_reference_f64(name):
    g, r = _run(name, np.float64)
    _check(g, r, "f64/", 1e-12, 1e-9)


@pytest.mark.parametrize("name", CASES)
def test_oracle_f32_matches_reference(name):
    g, r = _run(name, np.float32)
    # fp32 oracle vs fp64 reference: the 1e-5 fp32-check-mode budget of north_star
    _check(g, r, "f64/", 1e-5, 2e-4)
    _check(g, r, "f32/", 1e-5, 4e-4)


def test_make_batch_layout():
    ids, onehot, eps = vo.make_batch(5, 7)
    assert ids.shape == (7, 120) and ids.dtype == np.uint8
    assert onehot.shape == (7, 120, 35) and (onehot.sum(-1) == 1).all()
    assert (onehot.argmax(-1) == ids).all()
    lens = (ids != 0).sum(1)
    assert lens.min() >= 10 and lens.max() <= 110


def test_greedy_matches_step_argmax():
    P = vo.make_params(3, latent=8, hidden=16, layers=2)
    ids, onehot, eps = vo.make_batch(4, 3, latent=8)
    r = vo.config_b_step(P, onehot, eps, train=False, need_grads=False, layers=2)
    dec, _ = vo.greedy_decode(P, r["mu"], layers=2)
    assert (dec == r["argmax"]).all()


# ---- MOSES VAE oracle (oracle/moses_oracle.py) against fixtures from the reference's mosesvae.py ----
@pytest.mark.parametrize("name", ["moses_b6", "moses_b3_kl1"])
def test_moses_oracle_matches_reference(name):
    from oracle import moses_oracle as mo
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ps, bs, B = [int(v) for v in g["meta"]]
    klw = float(g["kl_weight"][0])
    P = mo.make_moses_params(ps, dtype=np.float64)
    seqs, eps, pad = mo.make_moses_batch(bs, B, dtype=np.float64)
    r = mo.moses_step(P, seqs, eps, pad, kl_weight=klw)
    assert abs(r["kl"] - g["f64/kl"]) <= 1e-11 * abs(g["f64/kl"])
    assert abs(r["recon"] - g["f64/recon"]) <= 1e-11 * abs(g["f64/recon"])
    np.testing.assert_allclose(r["z"], g["f64/z"], rtol=1e-10, atol=1e-12)
    # padded outputs: the reference applies decoder_fc to zero rows, i.e. the bias
    np.testing.assert_allclose(r["y"], g["f64/y"], rtol=1e-9, atol=1e-11)
    for k, gr in r["grads"].items():
        gn = float(g[f"f64/gnorm/{k}"])
        assert abs(np.sqrt((gr ** 2).sum()) - gn) <= 1e-9 * gn + 1e-14, k
        if f"f64/gfull/{k}" in g:
            np.testing.assert_allclose(gr, g[f"f64/gfull/{k}"], rtol=1e-8, atol=1e-12 + 1e-9 * gn)
        else:
            np.testing.assert_allclose(gr.reshape(-1)[g[f"f64/gidx/{k}"]], g[f"f64/gval/{k}"], rtol=1e-8,
                                       atol=1e-12 + 1e-9 * gn)


# ---- Config A oracle (oracle/cfga_oracle.py) against fixtures from the reference's models.py ----
@pytest.mark.parametrize("name", ["cfga_small_b3", "cfga_full_b2"])
def test_cfga_oracle_matches_reference(name):
    from oracle import cfga_oracle as ca
    g = np.load(os.path.join(GOLD, name + ".npz"))
    ps, bs, B = [int(v) for v in g["meta"]]
    eh, el, dh, dl, Z = [int(v) for v in g["cfg"]]
    P = ca.make_cfga_params(ps, dtype=np.float64, eh=eh, el=el, dh=dh, dl=dl, Z=Z)
    ids, onehot, eps = vo.make_batch(bs, B, latent=Z, dtype=np.float64)
    r = ca.cfga_step(P, ids.astype(np.int64), eps, el=el, dl=dl)
    assert abs(r["loss"] - float(g["loss"])) <= 1e-11 * abs(float(g["loss"]))
    np.testing.assert_allclose(r["probs"], g["probs"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r["mu"], g["mu"], rtol=1e-9, atol=1e-12)
    for k, gr in r["grads"].items():
        gn = float(g[f"gnorm/{k}"])
        assert abs(np.sqrt((gr ** 2).sum()) - gn) <= 1e-8 * gn + 1e-14, k
        if f"gfull/{k}" in g:
            np.testing.assert_allclose(gr, g[f"gfull/{k}"], rtol=1e-7, atol=1e-12 + 1e-9 * gn)
        else:
            np.testing.assert_allclose(gr.reshape(-1)[g[f"gidx/{k}"]], g[f"gval/{k}"], rtol=1e-7, atol=1e-12 + 1e-9 * gn)


def test_mosesfile_oracle_matches_reference():
    from oracle import moses_oracle as mo
    """oracle.moses_oracle.mosesfile_step against the fixture produced by the reference's mosesfile.VAE (--q_bidir)."""
    g = np.load(os.path.join(GOLD, "mosesfile_b5.npz"))
    ps, bs, B = [int(v) for v in g["meta"]]
    klw = float(g["kl_weight"][0])
    P = mo.make_mosesfile_params(ps, dtype=np.float64)
    seqs, eps, pad = mo.make_moses_batch(bs, B, d_z=128, dtype=np.float64)
    r = mo.mosesfile_step(P, seqs, eps, pad, kl_weight=klw)
    assert abs(r["kl"] - g["f64/kl"]) <= 1e-11 * abs(g["f64/kl"])
    assert abs(r["recon"] - g["f64/recon"]) <= 1e-11 * abs(g["f64/recon"])
    np.testing.assert_allclose(r["z"], g["f64/z"], rtol=1e-10, atol=1e-12)
    for k, gr in r["grads"].items():
        gn = float(g[f"f64/gnorm/{k}"])
        assert abs(np.sqrt((gr ** 2).sum()) - gn) <= 1e-9 * gn + 1e-14, k
        if f"f64/gfull/{k}" in g:
            np.testing.assert_allclose(gr, g[f"f64/gfull/{k}"], rtol=1e-8, atol=1e-12 + 1e-9 * gn)
        else:
            np.testing.assert_allclose(gr.reshape(-1)[g[f"f64/gidx/{k}"]], g[f"f64/gval/{k}"], rtol=1e-8, atol=1e-12 + 1e-9 * gn)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_binding_oracle_matches_reference(mode):
    """oracle.binding_oracle against the fixture produced by the reference's mosesvae.BindingModel."""
    from oracle import binding_oracle as bo
    g = np.load(os.path.join(GOLD, "binding_b12.npz"))
    ps, B, Z = [int(v) for v in g["meta"]]
    P, run = bo.make_binding_params(ps, Z, dtype=np.float64)
    r = bo.binding_step(P, run, g["z"], g["dout"], train=mode == "train")
    np.testing.assert_allclose(r["out"], g[f"{mode}/out"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(r["dz"], g[f"{mode}/dz"], rtol=1e-8, atol=1e-12)
    for k, v in r["grads"].items():
        np.testing.assert_allclose(v, g[f"{mode}/grad/{k}"], rtol=1e-8, atol=1e-11, err_msg=k)
    for k, v in r["running"].items():
        np.testing.assert_allclose(v, g[f"{mode}/{k}"], rtol=1e-10, atol=1e-12, err_msg=k)
