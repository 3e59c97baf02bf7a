from .quantizer import PerChannelAffineQuantizer

__all__ = ["PerChannelAffineQuantizer"]
