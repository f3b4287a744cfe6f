"""diffbir.sampler.sampler -> diffbir_b200.sampler.sampler."""
from diffbir_b200.sampler.sampler import Sampler  # noqa: F401
