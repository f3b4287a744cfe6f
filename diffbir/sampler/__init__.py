"""diffbir.sampler (reference sampler/__init__.py:1-4) -> diffbir_b200.sampler."""
from diffbir_b200.sampler import DDIMSampler, SpacedSampler, space_timesteps  # noqa: F401
from diffbir_b200.sampler.sampler import Sampler  # noqa: F401

from .._unsupported import unsupported

_alt = "use sampler 'spaced' or 'ddim' (same ControlLDM.forward; these loops are host code only)"
DPMSolverSampler = unsupported("DPMSolverSampler", "sampler/dpms_sampler.py:17-101", _alt)
EDMSampler = unsupported("EDMSampler", "sampler/edm_sampler.py:26-186", _alt)
