"""diffbir.sampler.spaced_sampler -> diffbir_b200.sampler.sampler."""
from diffbir_b200.sampler.sampler import SpacedSampler, space_timesteps  # noqa: F401
