"""On-disk formats of the reference, read and written unchanged:

  * weights: ``.pth`` = ``model.state_dict()`` (scripts/train.py:453,480,497; loaded at
    scripts/render_example.py:207), ``.npy`` / ``.npz`` = the official NeRF weight list consumed by
    ``NeRFMLP.load_from_numpy`` (model.py:83-127; file handling of render_example.py:166-205);
  * training checkpoints: ``{'model_state_dict', 'optimizer_state_dict', 'metrics'}``
    (scripts/train.py:471-475), resumed at :293-356.  ``optimizer_state_dict`` is
    ``torch.optim.Adam``'s, which FlatAdam reads and writes natively.

Host-side code only; after any load the model's bf16 weight image is rebuilt on next use."""
from __future__ import annotations

import numpy as np
import torch


def load_weights(model, path, map_location=None):
    """scripts/render_example.py:166-208: dispatch on the extension."""
    if str(path).endswith((".npy", ".npz")):
        weights = np.load(path, allow_pickle=True)
        if isinstance(weights, np.lib.npyio.NpzFile):
            weights = [weights[key] for key in weights.files]
        elif isinstance(weights, np.ndarray) and weights.dtype == object:
            weights = list(weights)
        model.load_from_numpy(weights)
    else:
        sd = torch.load(path, map_location=map_location or model.flat_params.device)
        if isinstance(sd, dict) and "model_state_dict" in sd:          # a training checkpoint was given
            sd = sd["model_state_dict"]
        model.load_state_dict(sd)
    return model


def save_weights(model, path):
    """``torch.save(model.state_dict(), path)`` (scripts/train.py:453): 24 fp32 tensors, reference keys."""
    torch.save({k: v.detach().clone() for k, v in model.state_dict().items()}, path)


def save_checkpoint(path, model, optimizer, metrics=None):
    """scripts/train.py:471-475 ('metrics_latest.pth')."""
    torch.save({"model_state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                "optimizer_state_dict": optimizer.state_dict(),
                "metrics": dict(metrics or {})}, path)


def load_checkpoint(path, model, optimizer=None, map_location=None):
    """scripts/train.py:293-356: restores model and optimizer, returns the metrics dict
    (``metrics.get('step', 0)`` is the step to resume from)."""
    ck = torch.load(path, map_location=map_location or model.flat_params.device)
    if "model_state_dict" in ck:
        model.load_state_dict(ck["model_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in ck:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return ck.get("metrics", {"step": ck.get("step", 0)})
