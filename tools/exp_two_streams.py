#!/usr/bin/env python
"""Experiment: one B = 8 plan vs two B = 4 plans whose step graphs replay concurrently on two streams (does the GroupNorm /
tail work of one half-batch hide under the tensor-bound convs of the other?).  Experiment tooling, not product code."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler  # noqa: E402
from clip_neural_image_conpression_b200.models import CLIPCondUNet  # noqa: E402


def make(dev):
    torch.manual_seed(0)
    net = CLIPCondUNet(z_dim=512, base=128, ch_mult=(1, 2, 2))
    with torch.no_grad():
        net.out.weight.mul_(0.1)
        net.out.bias.mul_(0.1)
    return net.to(dev).eval()


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    S, T = 256, 50
    parts = [int(a) for a in sys.argv[1:]] or [8, 4]
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(16, 512, generator=g), dim=-1).to(dev)
    x_T = torch.randn(16, 3, S, S, generator=g).to(dev)
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=0.0)
    for B in parts:
        n = 8 // B if B < 8 else 1
        nets = [make(dev) for _ in range(n)]
        streams = [torch.cuda.Stream() for _ in range(n)]
        def run():
            outs = []
            cur = torch.cuda.current_stream()
            for i, (net, st) in enumerate(zip(nets, streams)):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    outs.append(sampler.sample(net, z[i * B:(i + 1) * B], (B, 3, S, S), steps=T, x_T=x_T[i * B:(i + 1) * B]))
            for st in streams:
                cur.wait_stream(st)
            return outs
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        tot = B * n
        print(f"{n} stream(s) x batch {B}: {times[len(times) // 2]:8.2f} ms per decode of {tot} images -> {tot / times[len(times) // 2] * 1e3:6.2f} images/s")


if __name__ == "__main__":
    main()
