from .metrics import psnr, psnr_batch, ssim

__all__ = ["psnr", "psnr_batch", "ssim"]
