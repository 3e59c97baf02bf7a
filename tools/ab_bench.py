#!/usr/bin/env python
"""Single-process A/B of several builds / switch settings of libclpk.so on the config-2 plan (default UNet, 256 px, B = 8).

    python tools/ab_bench.py name=lib.so[,ENV=VAL,...] name2=... [--rounds 6] [--iters 5] [--batch 8]

Every variant gets its own plan (its library is loaded side by side with the others; environment switches are read at
plan creation).  Then the variants are profiled round-robin with clpk_plan_profile_steps (CUDA events around every
launch of `iters` eager DDIM steps), so all of them see the same thermal / power-cap state; medians over the rounds
are printed per kernel class.  "lib" may be `cur` for the in-tree build.  Experiment tooling, not product code."""
from __future__ import annotations

import ctypes as C
import os
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from clip_neural_image_conpression_b200 import _lib  # noqa: E402
from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler  # noqa: E402
from clip_neural_image_conpression_b200.models import CLIPCondUNet  # noqa: E402

CLASSES = ["conv_res", "conv_other", "groupnorm", "conv_in", "cond", "ddim"]


def load_lib(path: Path) -> C.CDLL:
    lib = C.CDLL(str(path))
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


def main() -> None:
    specs = [a for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    opts = {a.split("=")[0]: a.split("=")[1] for a in sys.argv[1:] if a.startswith("--") and "=" in a}
    rounds, iters, B = int(opts.get("--rounds", 6)), int(opts.get("--iters", 5)), int(opts.get("--batch", 8))
    S, T = int(opts.get("--size", 256)), 50
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cur = _lib.load()
    variants = []
    for spec in specs:
        name, rest = spec.split("=", 1)
        parts = rest.split(",")
        lib = cur if parts[0] == "cur" else load_lib((ROOT / parts[0]).resolve())
        env = dict(p.split("=", 1) for p in parts[1:])
        variants.append((name, lib, env))
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev)
    x_T = torch.randn(B, 3, S, S, generator=g).to(dev)
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=0.0)
    plans = []
    for name, lib, env in variants:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        _lib._lib = lib                      # the Python wrappers bind whatever library is current
        torch.manual_seed(0)
        net = CLIPCondUNet(z_dim=512, base=128, ch_mult=(1, 2, 2))
        with torch.no_grad():
            net.out.weight.mul_(0.1)
            net.out.bias.mul_(0.1)
        net = net.to(dev).eval()
        x = sampler.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)      # builds the plan + graph, warms up
        torch.cuda.synchronize()
        plans.append((name, lib, net, net.plan_for(B, S, S), x))
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    _lib._lib = cur
    ref = plans[0][4]
    for name, _, _, _, x in plans:
        d = (x - ref).double()
        print(f"{name:14s} finite={bool(torch.isfinite(x).all())} rel_l2_vs_first={float(d.norm() / ref.double().norm()):.2e}")
    stream = torch.cuda.current_stream()
    acc = {name: [[] for _ in CLASSES] for name, *_ in plans}
    for r in range(rounds):
        order = plans if r % 2 == 0 else plans[::-1]
        for name, lib, net, plan, _ in order:
            ms6, cnt6 = (C.c_float * 6)(), (C.c_int * 6)()
            rc = lib.clpk_plan_profile_steps(plan.handle, iters, ms6, cnt6, stream.cuda_stream)
            assert rc == 0, lib.clpk_last_error()
            for i in range(6):
                acc[name][i].append(ms6[i] / iters)
    # graph-replayed decode (what the product runs): CUDA events around one 50-step sample() per variant, round-robin
    graph_ms = {name: [] for name, *_ in plans}
    for r in range(max(rounds // 2, 3)):
        order = plans if r % 2 == 0 else plans[::-1]
        for name, lib, net, plan, _ in order:
            _lib._lib = lib
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sampler.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)
            e1.record(stream)
            torch.cuda.synchronize()
            graph_ms[name].append(e0.elapsed_time(e1) / T)
    _lib._lib = cur
    print(f"{'variant':14s} " + " ".join(f"{c:>10s}" for c in CLASSES) + f" {'total_ms':>10s} {'img/s':>8s} {'graph_ms':>9s} {'img/s':>8s}")
    for name, *_ in plans:
        med = [statistics.median(v) for v in acc[name]]
        tot = sum(med)
        gm = statistics.median(graph_ms[name])
        print(f"{name:14s} " + " ".join(f"{m:10.4f}" for m in med) + f" {tot:10.4f} {B / (tot * T) * 1e3:8.2f} {gm:9.4f} {B / (gm * T) * 1e3:8.2f}")


if __name__ == "__main__":
    main()
