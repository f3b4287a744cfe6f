from .dpm import DPMSolverSampler
from .edm import EDMSampler
from .sampler import DDIMSampler, EngineEval, Sampler, SpacedSampler, space_timesteps

__all__ = ["SpacedSampler", "DDIMSampler", "EDMSampler", "DPMSolverSampler", "Sampler", "EngineEval", "space_timesteps"]
