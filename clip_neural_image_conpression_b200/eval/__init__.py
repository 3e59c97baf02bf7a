from .metrics import psnr, psnr_batch, ssim
from .resample import bicubic_coeffs, load_original_device, resize_bicubic_u8, u8_hwc_to_float_chw

__all__ = ["psnr", "psnr_batch", "ssim", "bicubic_coeffs", "load_original_device", "resize_bicubic_u8", "u8_hwc_to_float_chw"]
