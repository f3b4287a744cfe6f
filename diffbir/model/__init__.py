"""diffbir.model (reference model/__init__.py:1-12) -> diffbir_b200.model."""
from diffbir_b200.model import ControlLDM, Diffusion, RRDBNet, SCUNet, SwinIR  # noqa: F401

from .._unsupported import unsupported
from . import config  # noqa: F401

_inside = "its forward lives inside diffbir.model.ControlLDM's kernel engine (diffbir_b200.engine)"
ControlledUnetModel = unsupported("ControlledUnetModel", "model/controlnet.py:16-47", _inside)
ControlNet = unsupported("ControlNet", "model/controlnet.py:50-328", _inside)
AutoencoderKL = unsupported("AutoencoderKL", "model/vae.py:562-582", _inside)
FrozenOpenCLIPEmbedder = unsupported("FrozenOpenCLIPEmbedder", "model/clip.py:9-61", _inside)
