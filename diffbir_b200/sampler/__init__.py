from .sampler import DDIMSampler, SpacedSampler, space_timesteps

__all__ = ["SpacedSampler", "DDIMSampler", "space_timesteps"]
