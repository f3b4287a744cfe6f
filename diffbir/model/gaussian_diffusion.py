"""diffbir.model.gaussian_diffusion -> diffbir_b200.model.gaussian_diffusion."""
from diffbir_b200.model.gaussian_diffusion import Diffusion  # noqa: F401
