"""B200-native decode hot path of the CLIP-feature neural image codec (drop-in for clip_feature_codec's decode side).

Sub-packages mirror the reference layout: models (CLIPCondUNet, ResBlock, FiLM, timestep_embedding), diffusion
(NoiseScheduler, DDIMSampler), io (.clp bitstreams), codecs (uint8 quantiser), eval (PSNR), cli (reconstruct_diffusion,
eval).  All compute runs in csrc/libclpk.so — hand-written sm_100a CUDA behind the C ABI of include/clpk.h.
"""
__version__ = "0.1.0"
