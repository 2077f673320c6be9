"""nerf_mlp_b200 -- B200-native (sm_100a) drop-in for the hot path of dgsmith7/nerf-mlp.

    from nerf_mlp_b200 import NeRFMLP, NeRFRenderer        # instead of `from nerfmlp import ...`

mirrors the reference package surface (nerfmlp/__init__.py:7-11) for the model and the renderer.
The dataset loader (nerfmlp/data.py) is host-side I/O outside the hot path and is not provided.
"""
from .model import NeRFMLP, PositionalEncoding
from .renderer import NeRFRenderer
from .optim import FlatAdam
from .train import TrainStep
from . import checkpoint, data, dist
from .data import DeviceRayDataset

__version__ = "1.0.0"
__all__ = ["NeRFMLP", "NeRFRenderer", "PositionalEncoding", "FlatAdam", "TrainStep", "DeviceRayDataset", "checkpoint", "data", "dist"]
