#!/usr/bin/env python
"""Small shapes through every round-2 kernel path, meant to run under compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Covers: fused head kernel, input-transform conv variants (pair + single CTA), shifted-statistics epilogue with residual and
16-bit-only residual output (transposed conv), affine fold, a whole forward + 3 graph-replayed DDIM steps of a small net at
128 px (row-slab levels, fused head, side-stream conditioning)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from clip_neural_image_conpression_b200 import ops  # noqa: E402
from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler  # noqa: E402
from clip_neural_image_conpression_b200.models import CLIPCondUNet  # noqa: E402

g = torch.Generator().manual_seed(0)
dev = "cuda"
# head kernel
x16 = torch.randn(2, 5, 256, 128, generator=g).half().to(dev)
sc, sh = torch.rand(2, 128, generator=g).to(dev) + 0.5, torch.randn(2, 128, generator=g).to(dev)
w, b = (torch.randn(3, 128, 3, 3, generator=g) * 0.03).to(dev), torch.randn(3, generator=g).to(dev)
y = ops.head_conv(x16, sc, sh, w, b)
y2 = ops.head_conv(x16[:, :, :40].contiguous(), sc, sh, w, b)
# input-transform convs
wt = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).to(dev)
bias = torch.randn(128, generator=g).to(dev)
o = ops.conv_igemm(x16, ops.pack_conv_weight(wt, 0), 0, 128, bias, in_affine=(sc, sh, True), gn_groups=8,
                   resid=torch.randn(2, 5, 256, 128, generator=g).to(dev), return_partial=True)
s2, h2 = ops.groupnorm_affine(o["gn_partial"], 2, o["gn_slots"], torch.ones(128, device=dev), torch.zeros(128, device=dev), 8)
o3 = ops.conv_igemm(x16, ops.pack_conv_weight(w, 0), 0, 3, b, in_affine=(sc, sh, False), want_f32=False, want_nchw=True)
# transposed conv with residual and 16-bit-only output
xt = torch.randn(1, 8, 8, 128, generator=g).half().to(dev)
wtt = (torch.randn(128, 64, 4, 4, generator=g) * 0.03).to(dev)
o4 = ops.conv_igemm(xt, ops.pack_conv_weight(wtt, 2), 2, 64, torch.zeros(64, device=dev), want_f32=False, want_op=True,
                    resid=torch.randn(1, 16, 16, 64, generator=g).to(dev), gn_groups=8)
# small net at 128 px: forward + graph loop
torch.manual_seed(0)
net = CLIPCondUNet(z_dim=512, base=64, ch_mult=(1, 2)).to(dev).eval()
z = torch.nn.functional.normalize(torch.randn(2, 512, generator=g), dim=-1).to(dev)
xT = torch.randn(2, 3, 128, 128, generator=g).to(dev)
eps = net(xT, z, torch.tensor([999, 10], device=dev))
x = DDIMSampler(NoiseScheduler(1000, "cosine", dev), 0.0).sample(net, z, (2, 3, 128, 128), steps=3, x_T=xT)
torch.cuda.synchronize()
ok = all(bool(torch.isfinite(t).all()) for t in (y, y2, o["f32"], s2, h2, o3["nchw"], o4["op"].float(), eps, x))
print("sanitize_small: finite =", ok)
sys.exit(0 if ok else 1)
