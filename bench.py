#!/usr/bin/env python
"""bench.py — DDIM-50 256 px decode throughput (images/s) of the B200 path, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

A "step" is one complete DDIM-50 decode of one batch (BASELINE.json configs[1]: default CLIPCondUNet base=128
ch_mult=(1,2,2), 256 px, eta=0, batch 8 per GPU): uint8 codes -> dequantise + L2 renorm -> 50 x (UNet + DDIM update,
replayed CUDA graph) -> clamp + uint8.  Weights are random-init (no checkpoints offline), data synthetic.
  value : images/s with the codes and x_T already resident in HBM (device timed, CUDA events, max over ranks);
  e2e   : same metric through the public host-buffer API: pinned-host codes + x_T -> H2D -> decode -> uint8 images D2H.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ddim50_256px_images_per_sec"
UNIT = "images/s"
ARCH = dict(z_dim=512, base=128, ch_mult=(1, 2, 2))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step (BASELINE configs[1]: 8)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--eta", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--operand", choices=["f16", "bf16"], default="f16", help="tensor-core operand format")
    ap.add_argument("--arch", choices=["default", "wide"], default="default",
                    help="default: base=128 ch_mult=(1,2,2) z=512 (BASELINE configs[1-3]); wide: base=192 ch_mult=(1,2,2,4) z=768 (configs[4], use --size 512 --batch 16)")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (one sample every 200 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001 — no nvidia-smi: report clocks as unknown
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(size: int, ddim_steps_sampled: int, repeats: int, warmup: int):
    """The reference algorithm (oracle port: identical ATen ops as PKG/models/unet.py + PKG/diffusion/ddim.py) on the
    host cores: B = 1 (the reference's eval loop is B = 1, eval.py:63), `ddim_steps_sampled` of the 50 DDIM steps per
    sample, extrapolated to 50.  Returns (images/s, ms per sample, cores)."""
    from oracle import codec_oracle as O   # checker / baseline only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(ARCH["z_dim"], ARCH["base"], ARCH["ch_mult"], seed=0, out_gain=0.1)
    tabs = O.scheduler_tables(1000, "cosine")
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(1, ARCH["z_dim"], generator=g), dim=-1)
    x = torch.randn(1, 3, size, size, generator=g)
    ts = O.ddim_timesteps(1000, 50)
    fn = lambda xx, zc, t: O.unet_forward(sd, ARCH["ch_mult"], xx, zc, t)  # noqa: E731

    def one_sample():
        xx = x
        with torch.no_grad():
            for i in range(ddim_steps_sampled):
                t = ts[i]
                eps = fn(xx, z, torch.full((1,), int(t), dtype=torch.long))
                a_s = tabs["alphas_cumprod_prev"][t] if i < 49 else torch.tensor(1.0)
                xx = O.ddim_update(xx, eps, tabs["alphas_cumprod"][t], a_s, 0.0)
        return xx

    for _ in range(warmup):
        one_sample()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one_sample()
        times.append(time.perf_counter() - t0)
    per_sample = sum(times) / len(times)
    per_ddim_step = per_sample / ddim_steps_sampled
    return 1.0 / (per_ddim_step * 50), per_sample * 1e3, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sampled = 2
    ips, ms, cores = cpu_reference_sample(args.size, sampled, max(args.steps, 1), max(args.warmup, 1))
    sample = (f"B=1 {args.size}px default UNet, {sampled} of 50 DDIM steps per sample on {cores} host threads, "
              f"extrapolated x{50 // sampled}; {args.steps} samples after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DDIM-50 {args.size}px default CLIPCondUNet base=128 ch_mult=(1,2,2), CPU fp32 (ATen)",
                   "batch": 1, "ddim_steps": 50, "eta": 0.0},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def conv_traffic_from_profile():
    """Bytes of DRAM traffic per ResBlock-conv launch from the committed ncu summary (None if it is missing)."""
    import re
    path = ROOT / "profiles" / "conv_traffic_r1.txt"
    try:
        m = re.search(r"= ([0-9.]+) MB per launch", path.read_text())
        return float(m.group(1)) * 1e6 if m else None
    except OSError:
        return None


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    from clip_neural_image_conpression_b200 import _lib, ops, parallel
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    from clip_neural_image_conpression_b200.pipeline import decode_codes

    rank, local, world = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, S, T = args.batch, args.size, args.ddim_steps

    # ---- synthetic store content: unit CLIP vectors -> uint8 codes; random-init weights, seed 0
    torch.manual_seed(0)
    net = CLIPCondUNet(**ARCH)
    with torch.no_grad():  # contractive head (protocol P-gamma) so the 50-step loop stays finite and non-saturated
        net.out.weight.mul_(0.1)
        net.out.bias.mul_(0.1)
    net.operand_dtype = torch.float16 if args.operand == "f16" else torch.bfloat16
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(5 + rank)
    Z = torch.nn.functional.normalize(torch.randn(B, ARCH["z_dim"], generator=g), dim=-1).to(dev)
    scale, zero = ops.quant_fit(Z)
    codes_dev = ops.quant_encode(Z, scale, zero)
    x_T_host = torch.randn(B, 3, S, S, generator=g).pin_memory()
    codes_host = codes_dev.cpu().pin_memory()
    out_host = torch.empty((B, S, S, 3), dtype=torch.uint8).pin_memory()
    metric_host = torch.empty(3, dtype=torch.float64).pin_memory()
    x_T_dev = x_T_host.to(dev)
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=args.eta)
    sampler.use_graph = not args.no_graph
    stream = torch.cuda.current_stream()

    # eval-style tail of every step (north-star: NCCL only AFTER the loop): PSNR + SSIM vs a synthetic "original" on the device,
    # then one gather of the uint8 reconstructions and one fp64 all-reduce of the metric sums
    target = torch.tanh(torch.randn(B, 3, S, S, generator=g)).to(dev)
    gathered = torch.empty((world * B, S, S, 3), dtype=torch.uint8, device=dev) if world > 1 else None

    def finish(x):
        u8 = ops.to_uint8_hwc(x)
        sq = ops.psnr_sqerr_u8(x, target)
        ss = ops.ssim_u8(x, target)
        sums = torch.stack([sq.sum().double(), ss.sum(), torch.tensor(float(B), dtype=torch.float64, device=dev)])
        if world > 1:
            torch.distributed.all_gather_into_tensor(gathered, u8)
            torch.distributed.all_reduce(sums)
        return u8, sums

    def step_resident():
        z = ops.dequant_l2norm(codes_dev, scale, zero)
        x = sampler.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T_dev)
        return finish(x)[0]

    def step_e2e():
        x = decode_codes(net, sampler, codes_host.numpy(), scale, zero, S, steps=T, batch=B,
                         x_T=x_T_host.to(dev, non_blocking=True))
        u8, sums = finish(x)
        out_host.copy_(u8, non_blocking=True)
        metric_host.copy_(sums, non_blocking=True)
        stream.synchronize()  # the user holds the images and the metric sums on the host when the call returns

    def timed(fn, k):
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        parallel.barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)  # ms, max over ranks

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = lib.clpk_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = int(lib.clpk_launch_count() - n0)
    clock_info = clocks.stop() if clocks else None
    value = world * B * args.steps / (ms_total / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- live per-kernel-class timing of the production kernels (eager launches bracketed by CUDA events)
    plan = net.plan_for(B, S, S)
    ms6, cnt6 = (C.c_float * 6)(), (C.c_int * 6)()
    prof_iters = 10
    _lib.check(lib.clpk_plan_profile_steps(plan.handle, prof_iters, ms6, cnt6, stream.cuda_stream), "profile")
    fa, fb, ge = C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.clpk_plan_work_breakdown(plan.handle, C.byref(fa), C.byref(fb), C.byref(ge)), "work")
    peaks = measured_peaks()
    cls = ["conv3x3_resblock_tcgen05", "conv_other_tcgen05", "groupnorm_silu", "conv_in", "conditioning", "ddim_update"]
    per_step_ms = {c: ms6[i] / prof_iters for i, c in enumerate(cls)}
    per_launch_ms = {c: (ms6[i] / cnt6[i] if cnt6[i] else None) for i, c in enumerate(cls)}
    conv_ms = per_step_ms[cls[0]]
    conv_tflops = fa.value / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    gn_ms = per_step_ms[cls[2]]
    gb = C.c_double()
    _lib.check(lib.clpk_plan_groupnorm_bytes(plan.handle, C.byref(gb)), "gn bytes")
    gn_bytes = gb.value  # algorithmic: input read once (2 B 16-bit / 4 B fp32) + 16-bit write per element
    gn_gbs = gn_bytes / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0
    n_conv = cnt6[0] // prof_iters
    roofline = {
        "bound": "tensor", "kernel": "conv_igemm_kernel (the 28 ResBlock 3x3 convs: row-slab CTA-pair variant at 256/128 px, CTA-pair k-block variant at 64/32 px)",
        "achieved": conv_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": conv_tflops / peaks["tf_sustained"], "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
        "traffic": conv_traffic_from_profile(),
        "traffic_note": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the same 28 launches of one DDIM step, from the committed ncu capture profiles/conv_traffic_r1.txt (not measured in this run); reads equal the algorithmic operand + residual bytes",
        "flops_per_launch": fa.value / max(n_conv, 1), "ms_per_launch": per_launch_ms[cls[0]],
        "launches_per_ddim_step": n_conv,
        "how": f"CUDA events around each of the {n_conv} launches of {prof_iters} eager DDIM steps right after the timed region, enqueued behind a stream-holding delay kernel (no host launch latency inside a pair) and with the cost of an empty event pair, calibrated in the same stream, subtracted; the per-class sum (unet_fwd_ms) reproduces the graph-replayed step time",
    }
    roofline_hbm = {
        "bound": "hbm", "kernel": "gn_apply_kernel (GroupNorm+SiLU, statistics fused into the producing conv)", "achieved": gn_gbs, "peak": peaks["hbm"],
        "unit": "GB/s", "frac": gn_gbs / peaks["hbm"], "algorithmic_bytes_per_ddim_step": gn_bytes,
        "gn_elements_per_ddim_step": ge.value, "bytes_per_element": gn_bytes / max(ge.value, 1.0),
        "note": "algorithmic traffic = input read once (16-bit: 2 B, fp32: 4 B) + 16-bit operand written once",
    }
    share = {c: per_step_ms[c] / sum(per_step_ms.values()) for c in cls}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        ips, ms, cores = cpu_reference_sample(S, 2, 2, 1)
        cpu_baseline = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"B=1 {S}px default UNet, 2 of 50 DDIM steps on {cores} host threads, extrapolated x25; 2 samples after 1 warm-up"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.operand, "data": "synthetic",
            "config": {"workload": f"DDIM-{T} {S}px {args.arch} CLIPCondUNet base={ARCH['base']} ch_mult={ARCH['ch_mult']}, eta={args.eta}, "
                                   f"batch {B} per GPU, {'CUDA-graph' if sampler.use_graph else 'eager'} step loop",
                       "batch_per_gpu": B, "global_batch": B * world, "ddim_steps": T, "z_dim": ARCH["z_dim"],
                       "weights": "random init (seed 0), out.* x0.1",
                       "precision": f"{args.operand} tensor-core operands, fp32 accumulation (TMEM), fp32 residual stream and GroupNorm statistics", "sharding": f"dp{world} by image, no collective in the DDIM loop; per step one all_gather of the uint8 reconstructions + one fp64 all_reduce of PSNR + SSIM sums (NCCL) when n_gpus > 1",
                       "l2": "per-step working set (>= 270 MB per activation tensor at batch 8) exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(codes_host.numel() + x_T_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() + metric_host.numel() * 8)},
            "gpu_launches": launches,
            "unet_fwd_ms": sum(per_step_ms[c] for c in cls[:5]),
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "kernel_ms_per_ddim_step": per_step_ms, "kernel_share": share,
            "cpu_baseline": cpu_baseline, "clocks": clock_info,
            "flops_per_image": float(plan.flops_per_forward) * T / B,
        }
        print(json.dumps(line), flush=True)
    parallel.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


def main():
    args = parse_args()
    if args.arch == "wide":
        ARCH.update(z_dim=768, base=192, ch_mult=(1, 2, 2, 4))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
