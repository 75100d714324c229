"""CPU-side checks: the C-ABI library loads, exports every symbol include/mvae_b200.h declares, and validates
arguments without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import molecular_vae_b200 as m
    return m._lib


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "mvae_b200.h")).read()
    names = set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", hdr))
    names -= {"mvae_stream_t"}
    L = _lib()
    missing = [n for n in sorted(names) if not hasattr(L.lib, n)]
    assert not missing, missing
    assert set(L.EXPORTED) <= names


def test_workspace_size_and_argument_validation():
    L = _lib()
    d = L.CfgBDesc(4096, 120, 35, 292, 501, 3, 435, L.PREC_BF16, 1, 120.0, 1.0)
    n16 = L.lib.mvae_cfgb_workspace_bytes(ctypes.byref(d))
    d32 = L.CfgBDesc(4096, 120, 35, 292, 501, 3, 435, L.PREC_FP32, 1, 120.0, 1.0)
    n32 = L.lib.mvae_cfgb_workspace_bytes(ctypes.byref(d32))
    assert 8e9 < n16 < 20e9 and n32 > 1.7 * n16
    bad = L.CfgBDesc(0, 120, 35, 292, 501, 3, 435, L.PREC_BF16, 1, 120.0, 1.0)
    assert L.lib.mvae_cfgb_workspace_bytes(ctypes.byref(bad)) == 0
    bad2 = L.CfgBDesc(8, 120, 35, 292, 501, 9, 435, L.PREC_BF16, 1, 120.0, 1.0)
    assert L.lib.mvae_cfgb_workspace_bytes(ctypes.byref(bad2)) == 0
    assert L.lib.mvae_strerror(-2).decode() == "workspace too small"
    with pytest.raises(L.MvaeError):
        L.check(-1)


def test_param_order_matches_reference_state_dict():
    import molecular_vae_b200 as m
    from oracle import vae_oracle as vo
    keys = m.param_order(3)
    assert keys == list(vo.config_b_shapes().keys())
    model = m.VAE(latent=292)
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k in sd.keys()] and set(sd.keys()) == set(keys)
    for k, shp in vo.config_b_shapes().items():
        assert tuple(sd[k].shape) == shp, k


def test_cfga_state_dict_matches_reference_layout():
    """models.MolecularVAE() exposes the reference's state_dict keys / shapes in C-ABI pointer order."""
    import molecular_vae_b200 as m
    from oracle import cfga_oracle as ca
    model = m.models.MolecularVAE()
    sd = model.state_dict()
    shapes = ca.cfga_shapes()
    assert list(sd.keys()) == list(shapes.keys()) == m.models.cfga_param_order(3, 4)
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == shp, k
    assert sum(p.numel() for p in model.parameters()) == 32285105
    L = m._lib
    d = L.CfgADesc(4096, 120, 35, 30, 72, 3, 292, 1024, 4, L.PREC_BF16, 120.0, 1e-2)
    assert 30e9 < L.lib.mvae_cfga_workspace_bytes(ctypes.byref(d)) < 80e9
    bad = L.CfgADesc(8, 120, 35, 30, 40, 3, 292, 1024, 4, L.PREC_BF16, 120.0, 1e-2)   # three k=18 convs do not fit
    assert L.lib.mvae_cfga_workspace_bytes(ctypes.byref(bad)) == 0


def test_no_cpu_fallback():
    import torch
    import molecular_vae_b200 as m
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(m._lib.MvaeError):
        m.CfgBEngine(4)
    with pytest.raises(m._lib.MvaeError):
        m.models.CfgAEngine(4)


def test_descriptor_validation_of_every_family():
    """Every *_workspace_bytes rejects malformed descriptions with 0 (no GPU needed); compute entry points validate their
    pointers before touching the device."""
    L = _lib()
    M = L.MosesDesc
    ok = M(64, 40, 34, 160, 256, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 0, 0, 0.0, 0)
    assert L.lib.mvae_moses_workspace_bytes(ctypes.byref(ok)) > 0
    for bad in (M(64, 40, 34, 160, 250, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 0, 0, 0.0, 0),      # q_hidden % 64
                M(64, 40, 300, 160, 256, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 0, 0, 0.0, 0),     # vocab > 256
                M(64, 1, 34, 160, 256, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 0, 0, 0.0, 0),       # max_len < 2
                M(64, 40, 34, 160, 256, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 0, 0, 1.5, 0)):     # dropout >= 1
        assert L.lib.mvae_moses_workspace_bytes(ctypes.byref(bad)) == 0
    bidir = M(64, 40, 34, 128, 256, 512, 3, 256, 32, L.PREC_BF16, 1.0, 1.0, 1, 1, 0.0, 0)
    assert L.lib.mvae_moses_workspace_bytes(ctypes.byref(bidir)) > L.lib.mvae_moses_workspace_bytes(ctypes.byref(ok)) * 0.9
    B = L.BindingDesc
    assert L.lib.mvae_binding_workspace_bytes(ctypes.byref(B(16, 128, 1, 1e-5, 0.1))) > 0
    assert L.lib.mvae_binding_workspace_bytes(ctypes.byref(B(0, 128, 1, 1e-5, 0.1))) == 0
    A = L.CfgADesc
    assert L.lib.mvae_cfga_workspace_bytes(ctypes.byref(A(8, 120, 35, 30, 72, 3, 292, 1000, 4, L.PREC_BF16, 120.0, 1e-2))) == 0  # dec_hidden % 64
    # null pointers are rejected before any CUDA call
    d = L.CfgBDesc(8, 120, 35, 292, 501, 3, 435, L.PREC_BF16, 1, 120.0, 1.0)
    rc = L.lib.mvae_cfgb_elbo_step(ctypes.byref(d), None, None, None, None, None, None, None, None, 0, None)
    assert rc in (-1, -2)
    rc = L.lib.mvae_ids_to_text(None, None, 4, 4, None, 1, None, -1, -1, 0, None, 0, None, None, None)
    assert rc == -1
