"""diffbir.sampler.dpms_sampler -> diffbir_b200.sampler.dpm."""
from diffbir_b200.sampler.dpm import DPMSolverSampler  # noqa: F401
