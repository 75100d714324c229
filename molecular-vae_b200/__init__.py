"""molecular-vae_b200: the ELBO hot path of aclyde11/molecular-VAE, hand-written for NVIDIA B200 (sm_100a).

Public surface (mirrors the reference's Python contract for this path):
  models2d.VAE / models2d.loss_function   drop-in for models2d.py + train.py:31-38 ("Config B")
  models.MolecularVAE / models.loss_function   drop-in for models.py as shipped ("Config A", LSTM encoder/decoder)
  engine.CfgBEngine                       direct access to the fused step / CUDA-graph replay
  _lib.lib                                the ctypes handle of the C ABI (include/mvae_b200.h)
Importing this package requires the built CUDA library; there is no CPU fallback.
"""
from . import _lib, checkpoint, ddp, engine, featurizer, models, models2d, mosesfile, mosesvae, optim, text  # noqa: F401
from .engine import CfgBEngine, param_order  # noqa: F401
from .models2d import VAE, loss_function  # noqa: F401
