"""Input featurisation: the drop-in OneHotFeaturizer against golden vectors written by the reference's own
class (tests/golden/make_golden_featurizer.py -> featurizer.npz), and its device path (`mvae_text_to_ids`) against the same."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "featurizer.npz")


def _load():
    g = np.load(GOLD)
    return g, [str(c) for c in g["charset"]], [str(s) for s in g["smiles"]]


def test_host_interface_matches_reference_vectors():
    import molecular_vae_b200 as m
    g, charset, smiles = _load()
    f = m.featurizer.OneHotFeaturizer(charset, 120)
    oh = f.featurize(smiles)
    assert oh.shape == g["onehot"].shape and oh.dtype.kind == "i" and (oh == g["onehot"]).all()
    assert [d[0] for d in f.one_hot_decode(oh)] == [str(s) for s in g["decoded"]]
    assert [f.decode_smiles_from_index(list(r)) for r in g["ids"]] == [str(s) for s in g["from_index"]]
    assert f.one_hot_array(3) == [int(i == 3) for i in range(len(charset))] and f.pad_smi("CC") == "CC" + " " * 118
    with pytest.raises(ValueError):
        f.featurize(["C?C"])                      # featurizer.py:17: charset.index raises


def test_host_interface_matches_reference_class_when_present():
    """In the build container the reference class itself is imported and compared on random strings."""
    import sys
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference not present on this box")
    import molecular_vae_b200 as m
    sys.path.insert(0, "/root/reference")
    try:
        import featurizer as ref
    finally:
        sys.path.remove("/root/reference")
    _, charset, _ = _load()
    rng = np.random.default_rng(3)
    smiles = ["".join(rng.choice(charset[1:], size=int(n))) for n in rng.integers(0, 60, size=40)]
    a, b = ref.OneHotFeaturizer(charset, 64), m.featurizer.OneHotFeaturizer(charset, 64)
    assert (a.featurize(smiles) == b.featurize(smiles)).all()
    assert a.one_hot_decode(a.featurize(smiles)) == b.one_hot_decode(b.featurize(smiles))


@pytest.mark.gpu
def test_device_featurisation_matches_reference_vectors():
    import molecular_vae_b200 as m
    g, charset, smiles = _load()
    f = m.featurizer.OneHotFeaturizer(charset, 120)
    ids = f.featurize_ids(smiles)
    assert ids.is_cuda and ids.dtype == torch.uint8 and tuple(ids.shape) == g["ids"].shape
    assert (ids.cpu().numpy() == g["ids"]).all()
    # a large ragged batch against the host path, and straight into the model
    rng = np.random.default_rng(5)
    big = ["".join(rng.choice(charset[1:], size=int(n))) for n in rng.integers(0, 121, size=3000)]
    got = f.featurize_ids(big).cpu().numpy()
    want = np.array([[charset.index(c) for c in s.ljust(120)] for s in big], dtype=np.uint8)
    assert (got == want).all()
    model = m.VAE(latent=292, precision="bf16").cuda()
    out = model.elbo_step(f.featurize_ids(big[:64]), torch.randn(64, 292, device="cuda"), use_graph=False)
    assert torch.isfinite(out).all()
    with pytest.raises(ValueError):
        f.featurize_ids(["CC", "C?"])
    with pytest.raises(ValueError):
        f.featurize_ids(["C" * 121])
