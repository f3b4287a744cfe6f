"""diffbir.sampler (reference sampler/__init__.py:1-4) -> diffbir_b200.sampler."""
from diffbir_b200.sampler import DDIMSampler, DPMSolverSampler, EDMSampler, SpacedSampler, space_timesteps  # noqa: F401
from diffbir_b200.sampler.sampler import Sampler  # noqa: F401
