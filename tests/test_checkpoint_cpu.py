"""Checkpoint compatibility: files in the reference's layouts load into the drop-in modules.
When /root/reference is present (build container) the files are written by the reference's OWN classes; on a box without
it the same layouts are produced from the drop-in modules (the state_dict contract is pinned by test_cabi_cpu.py)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


def _ref_module(name):
    if not os.path.isdir(REF):
        return None
    sys.path.insert(0, REF)
    try:
        return __import__(name)
    finally:
        sys.path.remove(REF)


def test_train_py_layout_with_dataparallel_prefix(tmp_path):
    import molecular_vae_b200 as m
    ref = _ref_module("models")
    torch.manual_seed(0)
    src = (ref.MolecularVAE() if ref else m.models.MolecularVAE())
    sd = src.state_dict()
    path = os.path.join(str(tmp_path), "save.pt")
    # train_distributed.py:145-151: saved from nn.DataParallel(model) -> `module.` prefix
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "optimizer_state_dict": {}, "epoch": 3,
                "charset": list("abc "), "max_len": 120, "lr": 1e-3}, path)
    dst = m.models.MolecularVAE()
    meta = m.checkpoint.load_reference_checkpoint(dst, path)
    assert meta["epoch"] == 3 and meta["max_len"] == 120 and meta["charset"] == list("abc ")
    for k, v in sd.items():
        assert torch.equal(dst.state_dict()[k], v), k
    # and back: a file we write is readable by the reference's own loader pattern (train_sample.py:16-19)
    out = os.path.join(str(tmp_path), "ours.pt")
    m.checkpoint.save_reference_checkpoint(dst, out, epoch=4, charset=list("abc "), max_len=120, lr=5e-4, latent_size=292)
    obj = torch.load(out, map_location="cpu", weights_only=False)
    assert set(obj) == {"model_state_dict", "optimizer_state_dict", "epoch", "charset", "max_len", "lr", "latent_size"}
    tgt = ref.MolecularVAE() if ref else m.models.MolecularVAE()
    tgt.load_state_dict(obj["model_state_dict"])


def test_moses_bare_state_dict_with_aliases(tmp_path):
    import molecular_vae_b200 as m
    from tests.test_gpu_moses import _Vocab
    ref = _ref_module("mosesvae")
    torch.manual_seed(1)
    src = ref.VAE(_Vocab()) if ref else m.mosesvae.VAE(_Vocab())
    sd = src.state_dict()                                        # includes encoder.N / decoder.N / vae.N aliases
    assert "vae.1.1.weight_hh_l0" in sd
    path = os.path.join(str(tmp_path), "trained_save.pt")
    torch.save({"module." + k: v for k, v in sd.items()}, path)  # mosesanalyize.py:168-175 expects this prefix
    dst = m.mosesvae.VAE(_Vocab())
    meta = m.checkpoint.load_reference_checkpoint(dst, path)
    assert meta == {}
    for k, v in sd.items():
        assert torch.equal(dst.state_dict()[k], v), k
    with pytest.raises(KeyError):
        m.checkpoint.load_reference_checkpoint(dst, {"bogus.weight": torch.zeros(1)})
