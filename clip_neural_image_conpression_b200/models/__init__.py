from .blocks import FiLM, ResBlock
from .unet import CLIPCondUNet, timestep_embedding

__all__ = ["CLIPCondUNet", "timestep_embedding", "FiLM", "ResBlock"]
