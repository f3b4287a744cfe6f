"""diffbir.sampler.ddim_sampler -> diffbir_b200.sampler.sampler."""
from diffbir_b200.sampler.sampler import DDIMSampler  # noqa: F401
