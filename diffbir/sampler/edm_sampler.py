"""diffbir.sampler.edm_sampler -> diffbir_b200.sampler.edm."""
from diffbir_b200.sampler.edm import EDMSampler  # noqa: F401
