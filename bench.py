#!/usr/bin/env python
"""bench.py — DDIM-50 256 px decode throughput (images/s) of the B200 path, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU implementation on the host cores

HEADLINE (`value`, `e2e`): a "step" is one complete DDIM-50 decode of one batch (BASELINE.json configs[1]: default
CLIPCondUNet base=128 ch_mult=(1,2,2), 256 px, eta=0, batch 8 per GPU): uint8 codes -> dequantise + L2 renorm ->
50 x (UNet + DDIM update, replayed CUDA graph) -> clamp + uint8 + PSNR/SSIM sums.  Weak scaling: 8 images per GPU.
  value : images/s with the codes and x_T already resident in HBM (device timed, CUDA events, max over ranks);
  e2e   : same metric through the public host-buffer API: pinned-host codes + x_T -> H2D -> decode -> uint8 images D2H.
EXTRA CONFIGURATIONS (key `extra_configs`, each measured once per run, never mixed into the headline; skip with --no-extras):
  bf16_operands      : the headline workload with bf16 tensor-core operands (the north-star's format) beside fp16;
  config3_store1024  : BASELINE configs[2] — a 1024-image synthetic store on disk (.clp + codec_meta.npz + manifest)
                       decoded through the eval path: read_bitstreams -> shard_bounds over the ranks -> micro-batches ->
                       uint8 + PSNR/SSIM -> one all_gather of the reconstructions + one all_reduce.  STRONG scaling:
                       total images fixed at every N (compare `seconds` across the per-N lines);
  config4_ddim250_b64: BASELINE configs[3] — DDIM-250, batch 64 per GPU, graph loop, eta = 1.0 (all-NaN like the reference)
                       and eta = 1e-3 (finite tensors: representative switching power);
  config5_wide_512px : BASELINE configs[4] — base=192 ch_mult=(1,2,2,4) z=768, 512 px, batch 16 per GPU.
Weights are random-init (no checkpoints offline), data synthetic.  Prints ONE JSON line on rank 0.
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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_configs measurements (configs 3/4/5, bf16)")
    ap.add_argument("--images", type=int, default=1024, help="store size of the config-3 strong-scaling measurement")
    ap.add_argument("--micro-batch", type=int, default=32,
                    help="micro-batch of the config-3 store decode (batch sweep on the final kernels, profiles/batch_sweep_r2.txt: 8 -> 61.4, 16 -> 64.0, 32 -> 66.5, 64 -> 67.0 images/s)")
    ap.add_argument("--only-store", action="store_true", help="run ONLY the config-3 store decode (value = its images/s, scaling strong)")
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
def load_staged_reference():
    """The unmodified reference package staged under oracle/_ref by oracle/stage_ref.py (None when it is absent)."""
    archive = ROOT / "oracle" / "_ref" / "clip_feature_codec.zip"
    if not archive.exists():
        return None
    if str(archive) not in sys.path:
        sys.path.insert(0, str(archive))   # zipimport: the archive holds the package exactly as the reference ships it
    try:
        import clip_feature_codec.diffusion.ddim as rd
        import clip_feature_codec.diffusion.scheduler as rs
        import clip_feature_codec.models.unet as ru
    except Exception:  # noqa: BLE001 — a broken staging copy must not kill the arm: fall back to the port
        return None
    return ru, rs, rd


def cpu_reference_sample(size: int, ddim_steps_sampled: int, repeats: int, warmup: int):
    """The reference's CPU implementation of the path on the host cores: B = 1 (the reference's eval loop is B = 1,
    eval.py:63), `ddim_steps_sampled` DDIM steps per sample, extrapolated to 50.  With oracle/_ref present this is the
    UNMODIFIED reference (CLIPCondUNet + DDIMSampler.sample with steps = ddim_steps_sampled: same per-step cost, the
    timestep values do not change the arithmetic); otherwise the oracle port (identical ATen ops).
    Returns (images/s, ms per sample, cores, kind)."""
    from oracle import codec_oracle as O   # checker / baseline only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict(ARCH["z_dim"], ARCH["base"], ARCH["ch_mult"], seed=0, out_gain=0.1)
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(1, ARCH["z_dim"], generator=g), dim=-1)
    x = torch.randn(1, 3, size, size, generator=g)
    staged = load_staged_reference()
    if staged is not None:
        ru, rs, rd = staged
        net = ru.CLIPCondUNet(z_dim=ARCH["z_dim"], base=ARCH["base"], ch_mult=ARCH["ch_mult"])
        net.load_state_dict(sd, strict=True)
        net.eval()
        sampler = rd.DDIMSampler(rs.NoiseScheduler(1000, "cosine", "cpu"), eta=0.0)
        kind = "reference"

        def one_sample():
            return sampler.sample(net, z, (1, 3, size, size), steps=ddim_steps_sampled, x_T=x)
    else:
        tabs = O.scheduler_tables(1000, "cosine")
        ts = O.ddim_timesteps(1000, 50)
        kind = "port"

        def one_sample():
            xx = x
            with torch.no_grad():
                for i in range(ddim_steps_sampled):
                    t = ts[i]
                    eps = O.unet_forward(sd, ARCH["ch_mult"], xx, z, torch.full((1,), int(t), dtype=torch.long))
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
    return 1.0 / (per_ddim_step * 50), per_sample * 1e3, cores, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sampled = 2
    ips, ms, cores, kind = cpu_reference_sample(args.size, sampled, max(args.steps, 1), max(args.warmup, 1))
    what = ("the unmodified reference (oracle/_ref: CLIPCondUNet + DDIMSampler.sample)" if kind == "reference"
            else "the oracle port of the reference algorithm")
    sample = (f"{what}: B=1 {args.size}px default UNet, {sampled} DDIM steps per sample on {cores} host threads, "
              f"extrapolated x{50 // sampled} to DDIM-50; {args.steps} samples after {args.warmup} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DDIM-50 {args.size}px default CLIPCondUNet base=128 ch_mult=(1,2,2), CPU fp32 (ATen)",
                   "batch": 1, "ddim_steps": 50, "eta": 0.0},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
CLASSES = ["conv3x3_resblock_tcgen05", "conv_other_tcgen05", "groupnorm_silu", "conv_in", "conditioning", "ddim_update"]


def build_net(arch, operand, dev):
    """Random-init CLIPCondUNet (seed 0) with a contractive head (protocol P-gamma) so long loops stay finite."""
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    torch.manual_seed(0)
    net = CLIPCondUNet(**arch)
    with torch.no_grad():
        net.out.weight.mul_(0.1)
        net.out.bias.mul_(0.1)
    net.operand_dtype = torch.float16 if operand == "f16" else torch.bfloat16
    return net.to(dev).eval()


def profile_plan(lib, _lib, plan, stream, iters):
    """Live per-kernel-class timing of the production kernels: CUDA events around every launch of `iters` eager DDIM steps
    (clpk_plan_profile_steps), plus the algorithmic work of the plan -> roofline figures."""
    ms6, cnt6 = (C.c_float * 6)(), (C.c_int * 6)()
    _lib.check(lib.clpk_plan_profile_steps(plan.handle, iters, ms6, cnt6, stream.cuda_stream), "profile")
    fa, fb, ge, gb = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.clpk_plan_work_breakdown(plan.handle, C.byref(fa), C.byref(fb), C.byref(ge)), "work")
    _lib.check(lib.clpk_plan_groupnorm_bytes(plan.handle, C.byref(gb)), "gn bytes")
    per_step = {c: ms6[i] / iters for i, c in enumerate(CLASSES)}
    per_launch = {c: (ms6[i] / cnt6[i] if cnt6[i] else None) for i, c in enumerate(CLASSES)}
    counts = {c: cnt6[i] // iters for i, c in enumerate(CLASSES)}
    conv_ms, gn_ms = per_step[CLASSES[0]], per_step[CLASSES[2]]
    return dict(per_step=per_step, per_launch=per_launch, counts=counts,
                conv_flops=fa.value, other_flops=fb.value, gn_elements=ge.value, gn_bytes=gb.value,
                conv_tflops=fa.value / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
                all_conv_tflops=(fa.value + fb.value) / ((conv_ms + per_step[CLASSES[1]]) * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
                gn_gbs=gb.value / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0)


def measure_dram_traffic(args):
    """Per-launch DRAM traffic of the ResBlock convs, read from the ncu table committed for THIS round
    (profiles/conv_traffic_r2.txt, regenerated by tools/summarize_conv_traffic.py from an ncu pass over this same bench
    command; its header names the commit).  ncu cannot run inside the timed process, so the figure is attached, not
    re-measured, and the key says which file it came from."""
    import re
    for name in ("conv_traffic_r2.txt", "conv_traffic_r1.txt"):
        path = ROOT / "profiles" / name
        try:
            m = re.search(r"= ([0-9.]+) MB per launch", path.read_text())
            if m:
                return float(m.group(1)) * 1e6, name
        except OSError:
            continue
    return None, None


def run_b200_arm(args):
    from clip_neural_image_conpression_b200 import _lib, ops, parallel
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    from clip_neural_image_conpression_b200.pipeline import decode_codes

    rank, local, world = parallel.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, S, T = args.batch, args.size, args.ddim_steps
    peaks = measured_peaks()
    stream = torch.cuda.current_stream()

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

    # ---- synthetic store content: unit CLIP vectors -> uint8 codes; random-init weights, seed 0
    net = build_net(ARCH, args.operand, dev)
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

    # eval-style tail of every step (north-star: NCCL only AFTER the loop): PSNR + SSIM vs a synthetic "original" on the device,
    # then one gather of the uint8 reconstructions and one fp64 all-reduce of the metric sums
    target = torch.tanh(torch.randn(B, 3, S, S, generator=g)).to(dev)
    gathered = torch.empty((world * B, S, S, 3), dtype=torch.uint8, device=dev) if world > 1 else None

    def finish(x, tgt=None, gat=None):
        tgt = target if tgt is None else tgt
        gat = gathered if gat is None else gat
        u8 = ops.to_uint8_hwc(x)
        sq = ops.psnr_sqerr_u8(x, tgt)
        ss = ops.ssim_u8(x, tgt)
        sums = torch.stack([sq.sum().double(), ss.sum(), torch.tensor(float(x.shape[0]), dtype=torch.float64, device=dev)])
        if world > 1:
            torch.distributed.all_gather_into_tensor(gat, u8)
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

    if args.only_store:
        cfg3 = extra_store(args, net, rank, world, dev, parallel, ops, timed_events=True)
        if rank == 0:
            line = {"metric": METRIC, "value": cfg3["images_per_sec"], "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                    "ms_per_step": cfg3["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": args.operand, "data": "synthetic", "config": {"workload": cfg3["workload"]}, "detail": cfg3}
            print(json.dumps(line), flush=True)
        parallel.barrier()
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
        return

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
    prof_iters = 10
    pr = profile_plan(lib, _lib, plan, stream, prof_iters)
    per_step_ms, per_launch_ms = pr["per_step"], pr["per_launch"]
    n_conv = pr["counts"][CLASSES[0]]
    traffic, traffic_file = measure_dram_traffic(args)
    roofline = {
        "bound": "tensor", "kernel": "conv_igemm_kernel (the 28 ResBlock 3x3 convs: row-slab CTA-pair variant at 256/128 px, CTA-pair k-block variant at 64/32 px)",
        "achieved": pr["conv_tflops"], "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": pr["conv_tflops"] / peaks["tf_sustained"], "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
        "traffic": traffic,
        "traffic_note": f"mean dram__bytes_read.sum + dram__bytes_write.sum per launch over the same {n_conv} launches of one DDIM step, from the committed ncu table profiles/{traffic_file} (ncu cannot run inside the timed process; the table's header names the command and commit)",
        "flops_per_launch": pr["conv_flops"] / max(n_conv, 1), "ms_per_launch": per_launch_ms[CLASSES[0]],
        "launches_per_ddim_step": n_conv, "all_convs_tflops": pr["all_conv_tflops"],
        "how": f"CUDA events around each of the {n_conv} launches of {prof_iters} eager DDIM steps right after the timed region, enqueued behind a stream-holding delay kernel (no host launch latency inside a pair) and with the cost of an empty event pair, calibrated in the same stream, subtracted; the per-class sum (unet_fwd_ms) reproduces the graph-replayed step time",
    }
    roofline_hbm = {
        "bound": "hbm", "kernel": "gn_apply_kernel (GroupNorm+SiLU, statistics fused into the producing conv)", "achieved": pr["gn_gbs"], "peak": peaks["hbm"],
        "unit": "GB/s", "frac": pr["gn_gbs"] / peaks["hbm"], "algorithmic_bytes_per_ddim_step": pr["gn_bytes"],
        "gn_elements_per_ddim_step": pr["gn_elements"], "bytes_per_element": pr["gn_bytes"] / max(pr["gn_elements"], 1.0),
        "launches_per_ddim_step": pr["counts"][CLASSES[2]],
        "note": "algorithmic traffic = input read once (16-bit: 2 B, fp32: 4 B) + 16-bit operand written once; GroupNorms that are applied inside the consuming conv (no stand-alone pass) are not part of this class",
    }
    share = {c: per_step_ms[c] / sum(per_step_ms.values()) for c in CLASSES}

    extras = {}
    if not args.no_extras and args.arch == "default" and (B, S, T) == (8, 256, 50):
        for name, fn in (("bf16_operands", lambda: extra_bf16(args, dev, rank, world, timed, finish, ops, sampler)),
                         ("config3_store1024", lambda: extra_store(args, net, rank, world, dev, parallel, ops)),
                         ("config4_ddim250_b64", lambda: extra_config4(args, net, dev, rank, world, lib, _lib, peaks, parallel)),
                         ("config5_wide_512px", lambda: extra_config5(args, dev, rank, world, lib, _lib, peaks, parallel))):
            try:
                extras[name] = fn()
            except Exception as e:  # noqa: BLE001 — an extra must never take the headline down; the failure is reported
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.synchronize()
            parallel.barrier()

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        ips, ms, cores, kind = cpu_reference_sample(S, 2, 2, 1)
        cpu_baseline = {"value": ips, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"B=1 {S}px default UNet, 2 of 50 DDIM steps on {cores} host threads, extrapolated x25; 2 samples after 1 warm-up"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.operand, "data": "synthetic",
            "config": {"workload": f"DDIM-{T} {S}px {args.arch} CLIPCondUNet base={ARCH['base']} ch_mult={ARCH['ch_mult']}, eta={args.eta}, "
                                   f"batch {B} per GPU (WEAK scaling: {B} images per GPU per step), {'CUDA-graph' if sampler.use_graph else 'eager'} step loop; "
                                   f"the 1024-image STRONG-scaling store decode (BASELINE configs[2]) is under extra_configs.config3_store1024",
                       "batch_per_gpu": B, "global_batch": B * world, "ddim_steps": T, "z_dim": ARCH["z_dim"],
                       "weights": "random init (seed 0), out.* x0.1",
                       "precision": f"{args.operand} tensor-core operands, fp32 accumulation (TMEM), fp32 GroupNorm statistics, residual stream "
                                    + ("fp16 at the levels with rows >= 128 px, fp32 below" if args.operand == "f16" else "fp32"), "sharding": f"dp{world} by image, no collective in the DDIM loop; per step one all_gather of the uint8 reconstructions + one fp64 all_reduce of PSNR + SSIM sums (NCCL) when n_gpus > 1",
                       "l2": (f"inputs larger than L2: every level-0 activation tensor is {B * S * S * ARCH['base'] * 2 / 1e6:.0f} MB at batch {B} "
                              f"(three of them live per ResBlock) vs the 126 MB L2; no flush "
                              + ("needed" if B * S * S * ARCH["base"] * 2 > 126e6 else "done although they FIT in L2 at this batch: not a headline configuration"))},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(codes_host.numel() + x_T_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() + metric_host.numel() * 8)},
            "gpu_launches": launches,
            "unet_fwd_ms": sum(per_step_ms[c] for c in CLASSES[:5]),
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "kernel_ms_per_ddim_step": per_step_ms, "kernel_share": share,
            "cpu_baseline": cpu_baseline, "clocks": clock_info,
            "flops_per_image": float(plan.flops_per_forward) * T / B,
            "extra_configs": extras,
        }
        print(json.dumps(line), flush=True)
    parallel.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------ extra configurations
def extra_bf16(args, dev, rank, world, timed, finish, ops, sampler):
    """The headline workload with bf16 tensor-core operands (the north-star's operand format) — 3 timed steps."""
    B, S, T = args.batch, args.size, args.ddim_steps
    net = build_net(ARCH, "bf16", dev)
    g = torch.Generator().manual_seed(50 + rank)
    z = torch.nn.functional.normalize(torch.randn(B, ARCH["z_dim"], generator=g), dim=-1).to(dev)
    x_T = torch.randn(B, 3, S, S, generator=g).to(dev)

    def step():
        finish(sampler.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T))

    for _ in range(3):
        step()
    ms = timed(step, 3)
    net.release_plans()
    return {"images_per_sec": world * B * 3 / (ms / 1e3), "ms_per_step": ms / 3, "dtype": "bf16", "steps": 3, "warmup": 3,
            "workload": "same as the headline (configs[1], weak), bf16 operands; parity of this variant: tests/test_gpu_fullsize.py"}


def extra_store(args, net, rank, world, dev, parallel, ops, timed_events=False):
    """BASELINE configs[2]: a synthetic store of --images .clp files decoded through the eval path, sharded by
    parallel.shard_bounds, STRONG scaling (total images fixed).  Timed once after a one-micro-batch warm-up, CUDA events on
    the stream (the first event is recorded before the host starts reading, so host-side zstd time is inside), max over ranks."""
    import shutil
    import tempfile

    from clip_neural_image_conpression_b200.codecs import PerChannelAffineQuantizer
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    from clip_neural_image_conpression_b200.io.bitstream import read_bitstreams, write_bitstream
    from clip_neural_image_conpression_b200.pipeline import decode_codes

    n, mb, S, T = args.images, args.micro_batch, args.size, args.ddim_steps
    store = [None]
    if rank == 0:
        d = Path(tempfile.mkdtemp(prefix="clpk_store_"))
        g = torch.Generator().manual_seed(77)
        Z = torch.nn.functional.normalize(torch.randn(n, ARCH["z_dim"], generator=g), dim=-1)
        qz = PerChannelAffineQuantizer(8).fit(Z.to(dev))
        codes = qz.encode_batch(Z.to(dev)) if hasattr(qz, "encode_batch") else ops.quant_encode(Z.to(dev), qz.scale, qz.zero)
        codes = codes.cpu().numpy()
        np.savez(d / "codec_meta.npz", scale=qz.scale.cpu().numpy(), zero=qz.zero.cpu().numpy(), dim=np.int32(ARCH["z_dim"]))
        manifest = []
        for i in range(n):
            write_bitstream(codes[i].tobytes(), ARCH["z_dim"], d / f"{i:05d}.clp")
            manifest.append({"image": f"synthetic:{i}", "bitstream": str(d / f"{i:05d}.clp")})
        (d / "manifest.json").write_text(json.dumps(manifest))
        store[0] = str(d)
    if world > 1:
        torch.distributed.broadcast_object_list(store, src=0)
    d = Path(store[0])
    manifest = json.loads((d / "manifest.json").read_text())
    meta = np.load(d / "codec_meta.npz")
    scale, zero = torch.from_numpy(meta["scale"]).to(dev), torch.from_numpy(meta["zero"]).to(dev)
    lo, hi = parallel.shard_bounds(n, rank, world)
    mine = manifest[lo:hi]
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=0.0)
    stream = torch.cuda.current_stream()
    torch.manual_seed(1000 + rank)
    # synthetic "originals" of this shard (the reference loads PNGs and BICUBIC-resizes them on the host, eval.py:66-67)
    orig = torch.tanh(torch.randn(hi - lo, 3, S, S, device=dev))

    def run(sub):
        q = read_bitstreams([Path(r["bitstream"]) for r in sub])
        x = decode_codes(net, sampler, q, scale, zero, S, steps=T, batch=mb)
        u8 = ops.to_uint8_hwc(x)
        sq = ops.psnr_sqerr_u8(x, orig[: len(sub)])
        ss = ops.ssim_u8(x, orig[: len(sub)])
        return u8, torch.stack([sq.sum().double(), ss.sum(), torch.tensor(float(len(sub)), dtype=torch.float64, device=dev)])

    run(mine[:mb])                                   # warm-up: plan + graph for the micro-batch shape
    torch.cuda.synchronize()
    parallel.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    u8, sums = run(mine)
    full = parallel.gather_shards(u8, n)             # one all_gather of the reconstructions (NCCL, after the loop)
    sums = parallel.reduce_sums(sums.tolist(), dev) if world > 1 else sums
    sums_host = sums.cpu()
    e1.record(stream)
    torch.cuda.synchronize()
    parallel.barrier()
    sec = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / 1e3
    assert full.shape[0] == n and int(sums_host[2]) == n
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    return {"workload": f"{n}-image synthetic .clp store (BASELINE configs[2]), DDIM-{T} {S}px default UNet, sharded contiguously over {world} GPU(s), micro-batch {mb}, STRONG scaling (total images fixed)",
            "images": n, "n_gpus": world, "images_per_gpu": hi - lo, "micro_batch": mb, "seconds": sec, "images_per_sec": n / sec,
            "scaling": "strong", "includes": "host .clp read + zstd (thread pool), H2D of the codes, dequant + L2, DDIM loop (device-drawn x_T), uint8 + PSNR + SSIM on the device, one all_gather of the uint8 images, one all_reduce + D2H of the metric sums",
            "timing": "CUDA events on the stream around the whole pass (recorded before the host starts reading), max over ranks"}


def extra_config4(args, net, dev, rank, world, lib, _lib, peaks, parallel):
    """BASELINE configs[3]: default UNet, 256 px, DDIM-250, batch 64 per GPU, graph loop with in-graph Philox noise.
    eta = 1.0 reproduces the reference's all-NaN output; eta = 1e-3 keeps the tensors finite."""
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    B, S, T = 64, args.size, 250
    g = torch.Generator().manual_seed(60 + rank)
    z = torch.nn.functional.normalize(torch.randn(B, ARCH["z_dim"], generator=g), dim=-1).to(dev)
    x_T = torch.randn(B, 3, S, S, generator=g).to(dev)
    stream = torch.cuda.current_stream()
    out = {"workload": f"DDIM-{T} {S}px default UNet, batch {B} per GPU (weak), CUDA-graph step loop, in-graph Philox noise",
           "batch_per_gpu": B, "ddim_steps": T}
    for tag, eta, warm in (("eta_1e-3", 1e-3, True), ("eta_1.0", 1.0, False)):
        s = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=eta)
        if warm:
            s.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)
        torch.cuda.synchronize()
        parallel.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        x = s.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)
        e1.record(stream)
        torch.cuda.synchronize()
        sec = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / 1e3
        out[tag] = {"seconds_per_decode": sec, "images_per_sec": world * B / sec,
                    "nan_fraction": float(torch.isnan(x).float().mean())}
    pr = profile_plan(lib, _lib, net.plan_for(B, S, S), stream, 3)
    out["roofline"] = {"bound": "tensor", "kernel": "conv_igemm_kernel (28 ResBlock convs, batch 64)", "achieved": pr["conv_tflops"],
                       "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": pr["conv_tflops"] / peaks["tf_sustained"],
                       "note": "profiled on the eta = 1.0 state (x is NaN after step 0: low switching power, optimistic clocks); the eta_1e-3 throughput above is the representative figure"}
    out["kernel_ms_per_ddim_step"] = pr["per_step"]
    net.release_plans()
    return out


def extra_config5(args, dev, rank, world, lib, _lib, peaks, parallel):
    """BASELINE configs[4]: wide UNet base=192 ch_mult=(1,2,2,4), 768-d conditioning, 512 px, DDIM-50, batch 16 per GPU."""
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    arch = dict(z_dim=768, base=192, ch_mult=(1, 2, 2, 4))
    B, S, T = 16, 512, 50
    net = build_net(arch, args.operand, dev)
    g = torch.Generator().manual_seed(70 + rank)
    z = torch.nn.functional.normalize(torch.randn(B, 768, generator=g), dim=-1).to(dev)
    x_T = torch.randn(B, 3, S, S, generator=g).to(dev)
    s = DDIMSampler(NoiseScheduler(1000, "cosine", dev), eta=0.0)
    stream = torch.cuda.current_stream()
    s.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)
    torch.cuda.synchronize()
    parallel.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    x = s.sample(net, z, (B, 3, S, S), steps=T, x_T=x_T)
    e1.record(stream)
    torch.cuda.synchronize()
    sec = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / 1e3
    plan = net.plan_for(B, S, S)
    pr = profile_plan(lib, _lib, plan, stream, 2)
    out = {"workload": f"DDIM-{T} {S}px wide UNet base=192 ch_mult=(1,2,2,4) z=768, batch {B} per GPU (weak), CUDA-graph step loop",
           "batch_per_gpu": B, "seconds_per_decode": sec, "images_per_sec": world * B / sec, "finite": bool(torch.isfinite(x).all()),
           "flops_per_image": float(plan.flops_per_forward) * T / B, "plan_device_bytes": plan.device_bytes,
           "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (36 ResBlock convs of the wide net)", "achieved": pr["conv_tflops"],
                        "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": pr["conv_tflops"] / peaks["tf_sustained"]},
           "kernel_ms_per_ddim_step": pr["per_step"]}
    net.release_plans()
    del net
    torch.cuda.empty_cache()
    return out


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
