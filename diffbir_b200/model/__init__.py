from .bsrnet import RRDBNet
from .cldm import ControlLDM
from .gaussian_diffusion import Diffusion
from .scunet import SCUNet
from .swinir import SwinIR

__all__ = ["ControlLDM", "Diffusion", "SwinIR", "RRDBNet", "SCUNet"]
