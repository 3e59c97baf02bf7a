from .ddim import DDIMSampler
from .scheduler import NoiseScheduler

__all__ = ["DDIMSampler", "NoiseScheduler"]
