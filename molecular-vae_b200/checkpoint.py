"""Checkpoint I/O in the reference's own file layouts, so that files written by the reference's
training scripts load into the B200 drop-in modules and vice versa.

Layouts handled:
  * train.py:170-177       dict {model_state_dict, optimizer_state_dict, epoch, charset, max_len, lr, latent_size}
  * train_distributed.py:145-151  same dict without latent_size; the keys carry DataParallel's `module.` prefix
  * moses_train_distrib*.py:343-345 / :534   bare state_dict (possibly `module.`-prefixed, stripped by
    mosesanalyize.py:168-175), incl. the ModuleList aliases `encoder.N.*`, `decoder.N.*`, `vae.N.*` of mosesvae.py:90-105
Host-side only (torch.load / torch.save); no device work."""
from collections import OrderedDict

import torch

META_KEYS = ("optimizer_state_dict", "epoch", "charset", "max_len", "lr", "latent_size")


def strip_module_prefix(state_dict):
    """mosesanalyize.py:171-173: drop the `module.` that nn.DataParallel / DistributedDataParallel prepend."""
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def load_reference_checkpoint(model, source, map_location="cpu", strict=True, allow_pickle=False):
    """Load a checkpoint written by the reference (path or already-loaded object) into `model` (a drop-in module of this
    package or the reference's own class).  Returns the metadata dict of the train.py layout (empty for bare state_dicts).
    The reference layouts hold only tensors, dicts, lists, ints, floats and strings, so files are read with
    `weights_only=True` (no arbitrary pickle code); `allow_pickle=True` is the explicit opt-in for anything else."""
    obj = (torch.load(source, map_location=map_location, weights_only=not allow_pickle)
           if isinstance(source, (str, bytes)) or hasattr(source, "read") else source)
    meta = {}
    if isinstance(obj, dict) and "model_state_dict" in obj:
        meta = {k: obj[k] for k in META_KEYS if k in obj}
        sd = obj["model_state_dict"]
    else:
        sd = obj
    sd = strip_module_prefix(sd)
    own = model.state_dict()
    # checkpoints of wrapped / aliased modules may hold extra alias keys; keep what the target knows
    unknown = [k for k in sd if k not in own]
    if unknown and strict:
        raise KeyError(f"checkpoint keys not present in the model: {unknown[:5]}{'...' if len(unknown) > 5 else ''}")
    model.load_state_dict(OrderedDict((k, v) for k, v in sd.items() if k in own), strict=strict)
    return meta


def save_reference_checkpoint(model, path, optimizer=None, data_parallel_prefix=False, bare=False, **meta):
    """Write `model` in the reference's layout: train.py:170-177 dict (default; pass epoch=, charset=, max_len=, lr=,
    latent_size=), with `module.`-prefixed keys when data_parallel_prefix (train_distributed.py:145), or the bare
    state_dict of the moses scripts (bare=True)."""
    sd = model.state_dict()
    if data_parallel_prefix:
        sd = OrderedDict(("module." + k, v) for k, v in sd.items())
    if bare:
        torch.save(sd, path)
        return
    obj = {"model_state_dict": sd, "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {}}
    obj.update({k: v for k, v in meta.items() if k in META_KEYS})
    torch.save(obj, path)
