"""GPU parity of the plan-level path (whole CLIPCondUNet forward, DDIM loop, CUDA graph) against the golden outputs of
the unmodified reference and the oracle.  Tolerances are the north-star's: per-step epsilon <= 1e-2 relative L2
(teacher forced), final reconstruction >= 40 dB PSNR vs the reference's fp32 output (contractive weights)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(z_dim=512, base=32, ch_mult=(1, 2))
MID = dict(z_dim=512, base=64, ch_mult=(1, 2))
EPS_TOL = 1e-2
PSNR_BAR = 40.0


def make_net(oracle, cfg, seed, out_gain=1.0):
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    net = CLIPCondUNet(z_dim=cfg["z_dim"], base=cfg["base"], ch_mult=cfg["ch_mult"])
    sd = oracle.make_state_dict(cfg["z_dim"], cfg["base"], cfg["ch_mult"], seed=seed, out_gain=out_gain)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), sd


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def test_reference_shape_test():
    """Mirror of the reference's tests/test_unet.py:7-13 on the CUDA path."""
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    net = CLIPCondUNet(z_dim=512, base=64, ch_mult=(1, 2), img_ch=3).cuda()
    x, z, t = torch.randn(2, 3, 64, 64).cuda(), torch.randn(2, 512).cuda(), torch.randint(0, 1000, (2,)).cuda()
    y = net(x, z, t)
    assert y.shape == x.shape and torch.isfinite(y).all()


@pytest.mark.parametrize("name,cfg", [("tiny", TINY), ("mid", MID)])
def test_unet_forward_vs_reference_golden(oracle, golden, name, cfg):
    g = golden("unet_forward")
    net, _ = make_net(oracle, cfg, seed=11)
    eps = net(cu(g[f"{name}.x"]), cu(g[f"{name}.z"]), cu(g[f"{name}.t"])).cpu()
    ref = torch.from_numpy(g[f"{name}.eps"])
    for b in range(ref.shape[0]):  # per sample: t = 999 and t = 37
        assert oracle.rel_l2(eps[b], ref[b]) < EPS_TOL, (name, b, oracle.rel_l2(eps[b], ref[b]))


def test_bf16_operand_variant(oracle, golden):
    """bf16 operands (selectable) meet the epsilon bar on >= 64-channel networks; fp16 (default) is ~8x tighter."""
    g = golden("unet_forward")
    net, _ = make_net(oracle, MID, seed=11)
    ref = torch.from_numpy(g["mid.eps"])
    e16 = oracle.rel_l2(net(cu(g["mid.x"]), cu(g["mid.z"]), cu(g["mid.t"])).cpu(), ref)
    net.operand_dtype = torch.bfloat16
    eb = oracle.rel_l2(net(cu(g["mid.x"]), cu(g["mid.z"]), cu(g["mid.t"])).cpu(), ref)
    assert e16 < 2e-3 and eb < EPS_TOL and e16 < eb, (e16, eb)


def test_unet_forward_is_deterministic_and_batch_independent(oracle):
    net, _ = make_net(oracle, TINY, seed=3)
    g = torch.Generator().manual_seed(1)
    x, z = torch.randn(4, 3, 32, 32, generator=g).cuda(), torch.randn(4, 512, generator=g).cuda()
    t = torch.tensor([999, 500, 20, 0]).cuda()
    a, b = net(x, z, t), net(x, z, t)
    assert torch.equal(a, b)
    solo = torch.cat([net(x[i:i + 1], z[i:i + 1], t[i:i + 1]) for i in range(4)])
    # images never interact (GroupNorm / FiLM are per sample); only the GN partial-sum split depends on the batch
    assert oracle.rel_l2(solo, a) < 1e-5


def test_state_dict_reload_rebuilds_plan(oracle):
    net, sd = make_net(oracle, TINY, seed=3)
    g = torch.Generator().manual_seed(1)
    x, z, t = torch.randn(1, 3, 32, 32, generator=g).cuda(), torch.randn(1, 512, generator=g).cuda(), torch.tensor([7]).cuda()
    a = net(x, z, t)
    with torch.no_grad():
        net.out.weight.mul_(0.5)
        net.out.bias.mul_(0.5)
    b = net(x, z, t)
    assert oracle.rel_l2(b, 0.5 * a) < 1e-5


def _sampler(eta):
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    return DDIMSampler(NoiseScheduler(1000, "cosine", "cuda"), eta=eta)


def test_scheduler_tables_on_device_bit_exact(golden):
    from clip_neural_image_conpression_b200.diffusion import NoiseScheduler
    g = golden("scheduler")
    for sch in ("cosine", "linear"):
        s = NoiseScheduler(1000, sch, "cuda")
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "posterior_variance"):
            assert getattr(s, k).is_cuda and np.array_equal(getattr(s, k).cpu().numpy(), g[f"{sch}.{k}"])


def test_ddim_config1_vs_reference_golden(oracle, golden):
    """BASELINE configs[0]: base=32 ch_mult=(1,2) 64 px, batch 2, DDIM 10 steps — closed loop vs the reference."""
    g = golden("ddim")
    z, x_T = cu(g["z"]), cu(g["x_T"])
    net, _ = make_net(oracle, TINY, seed=0, out_gain=0.1)
    x = _sampler(0.0).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T).cpu()
    assert oracle.psnr_float(x, torch.from_numpy(g["pg.eta0.x"])) >= PSNR_BAR
    x50 = _sampler(0.0).sample(net, z, (2, 3, 64, 64), steps=50, x_T=x_T).cpu()
    assert oracle.psnr_float(x50, torch.from_numpy(g["pg.eta0.steps50.x"])) >= PSNR_BAR


def test_ddim_stochastic_path_and_nan_pattern(oracle, golden):
    g = golden("ddim")
    z, x_T = cu(g["z"]), cu(g["x_T"])
    net, _ = make_net(oracle, TINY, seed=0, out_gain=0.1)
    torch.manual_seed(7)
    noise = torch.stack([torch.randn(2, 3, 64, 64) for _ in range(10)]).cuda()  # the reference's global-RNG draws
    x = _sampler(1e-3).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T, noise=noise).cpu()
    assert oracle.psnr_float(x, torch.from_numpy(g["pg.eta1e-3.x"])) >= PSNR_BAR
    # eta = 1.0: the reference's sqrt(a_s - sigma^2) is NaN from step 0 on -> all-NaN output
    xn = _sampler(1.0).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T, noise=noise)
    assert float(torch.isnan(xn).float().mean()) == float(g["pg.eta1.nan_fraction"]) == 1.0
    # in-kernel Philox noise.  Pinned key: finite, reproducible, key-sensitive.
    s = _sampler(1e-3)
    s.seed = 0
    a = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    b = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    s.seed = 1
    c = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    assert torch.isfinite(a).all() and torch.equal(a, b) and not torch.equal(a, c)
    # Default (seed None): a fresh key per call from torch's generator, like the reference's torch.randn_like draws
    # (ddim.py:44-45) — calls differ, torch.manual_seed makes the sequence reproducible.
    s = _sampler(1e-3)
    torch.manual_seed(123)
    d1, d2 = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T), s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    torch.manual_seed(123)
    d3 = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    assert not torch.equal(d1, d2) and torch.equal(d1, d3)


def test_ddim_teacher_forced_epsilon(oracle, golden):
    """Open-loop per-step epsilon parity on plain (undamped) random weights: the oracle's own x_t trajectory is fed to
    the CUDA UNet at every sampled step; relative L2 of eps must stay under 1e-2 (north-star)."""
    g = golden("ddim")
    z, x_T = torch.from_numpy(g["z"]), torch.from_numpy(g["x_T"])
    net, sd = make_net(oracle, TINY, seed=0)
    tabs = oracle.scheduler_tables(1000, "cosine")
    tr = {}
    with torch.no_grad():
        oracle.ddim_sample(lambda x, zc, t: oracle.unet_forward(sd, (1, 2), x, zc, t), tabs, z, x_T, steps=10, trace=tr)
    ts = oracle.ddim_timesteps(1000, 10)
    worst = 0.0
    for i in range(10):
        t_b = torch.full((2,), int(ts[i]), dtype=torch.long).cuda()
        eps = net(tr["x"][i].cuda(), z.cuda(), t_b).cpu()
        worst = max(worst, oracle.rel_l2(eps, tr["eps"][i]))
    assert worst < EPS_TOL, worst


def test_graph_replay_equals_eager_launches_and_trace(oracle, golden):
    g = golden("ddim")
    z, x_T = cu(g["z"]), cu(g["x_T"])
    net, _ = make_net(oracle, TINY, seed=0, out_gain=0.1)
    s = _sampler(0.0)
    tr = {}
    a = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T, trace=tr)
    s.use_graph = False
    b = s.sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    assert torch.equal(a, b)                                   # same kernels, same order -> bitwise identical
    assert tr["x"].shape == (10, 2, 3, 64, 64) and torch.equal(tr["x"][0], x_T)
    # the traced eps of step 0 is what forward() returns for x_T at t = 999
    e0 = net(x_T, z, torch.full((2,), 999, device="cuda"))
    assert oracle.rel_l2(tr["eps"][0], e0) < 1e-5
    # generic-callable path (any eps model) walks the same trajectory
    c = s.sample(lambda x, zc, t: net(x, zc, t), z, (2, 3, 64, 64), steps=10, x_T=x_T)
    assert oracle.psnr_float(c.cpu(), a.cpu()) > 60.0


def test_default_architecture_full_size_properties(oracle):
    """BASELINE configs[1] shape (base=128, ch_mult=(1,2,2), 256 px, batch 8): too slow for the CPU oracle, so the
    full-size run is checked through size-independent properties + the SAME oracle code executed on the GPU in fp32."""
    cfg = dict(z_dim=512, base=128, ch_mult=(1, 2, 2))
    net, sd = make_net(oracle, cfg, seed=21, out_gain=0.1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 3, 256, 256, generator=g).cuda()
    z = torch.nn.functional.normalize(torch.randn(8, 512, generator=g), dim=-1).cuda()
    t = torch.tensor([999, 978, 795, 489, 250, 40, 20, 0]).cuda()
    eps = net(x, z, t)
    assert torch.isfinite(eps).all()
    assert torch.equal(eps, net(x, z, t))                                             # idempotent / deterministic
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4]).cuda()
    assert oracle.rel_l2(net(x[perm], z[perm], t[perm]), eps[perm]) < 1e-5            # equivariant to batch order
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_cu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = oracle.unet_forward(sd_cu, cfg["ch_mult"], x, z, t)                     # oracle code, fp32, on the GPU
    for b in range(8):
        assert oracle.rel_l2(eps[b], ref[b]) < EPS_TOL, (b, oracle.rel_l2(eps[b], ref[b]))
    # 10-step closed loop at full size against the same oracle loop on the GPU
    s = _sampler(0.0)
    x_T = torch.randn(8, 3, 256, 256, generator=g).cuda()
    out = s.sample(net, z, (8, 3, 256, 256), steps=10, x_T=x_T)
    tabs = oracle.scheduler_tables(1000, "cosine")
    with torch.no_grad():
        ref = oracle.ddim_sample(lambda xx, zc, tt: oracle.unet_forward(sd_cu, cfg["ch_mult"], xx, zc, tt), tabs, z, x_T, steps=10)
    assert oracle.psnr_float(out, ref) >= PSNR_BAR


def test_wide_architecture_reduced_resolution(oracle):
    """BASELINE configs[4] architecture (base=192, ch_mult=(1,2,2,4), 768-d CLIP ViT-L/14 vectors; channels
    192/384/768/3072) at 32 px so the fp32 oracle stays cheap: exercises N = 192 tiles, GroupNorm groups of 24/48
    channels (separate statistics pass), 96/384 channels (fused), 12 channel tiles per conv, 2x2-pixel bottom level."""
    cfg = dict(z_dim=768, base=192, ch_mult=(1, 2, 2, 4))
    net, sd = make_net(oracle, cfg, seed=31, out_gain=0.1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 32, 32, generator=g).cuda()
    z = torch.nn.functional.normalize(torch.randn(2, 768, generator=g), dim=-1).cuda()
    t = torch.tensor([999, 123]).cuda()
    eps = net(x, z, t)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_cu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = oracle.unet_forward(sd_cu, cfg["ch_mult"], x, z, t)
    for b in range(2):
        assert oracle.rel_l2(eps[b], ref[b]) < EPS_TOL, (b, oracle.rel_l2(eps[b], ref[b]))
    del net, sd_cu
    torch.cuda.empty_cache()


def test_ddim_250_steps_stochastic_graph_loop(oracle, golden):
    """BASELINE configs[3] loop shape (250 steps, eta = 1.0, graph replay) on the small net: the reference's eta = 1
    output is all-NaN and so is ours; eta = 1e-3 walks the same 250 unique timesteps with finite values and
    matches the oracle fed with the same pre-drawn noise."""
    g = golden("ddim")
    z, x_T = cu(g["z"]), cu(g["x_T"])
    net, sd = make_net(oracle, TINY, seed=0, out_gain=0.1)
    xn = _sampler(1.0).sample(net, z, (2, 3, 64, 64), steps=250, x_T=x_T)
    assert bool(torch.isnan(xn).all())
    noise = torch.randn(250, 2, 3, 64, 64, generator=torch.Generator().manual_seed(8))
    tr = {}
    x = _sampler(1e-3).sample(net, z, (2, 3, 64, 64), steps=250, x_T=x_T, noise=noise.cuda(), trace=tr)
    assert torch.isfinite(x).all()
    tabs = oracle.scheduler_tables(1000, "cosine")
    ref = {}
    with torch.no_grad():
        oracle.ddim_sample(lambda xx, zc, t: oracle.unet_forward(sd, (1, 2), xx, zc, t), tabs, z.cpu(), x_T.cpu(),
                           steps=250, eta=1e-3, noise=noise, trace=ref)
    # a 250-step closed loop amplifies any epsilon difference chaotically (the reference's own fp64-vs-fp32 runs diverge,
    # SURVEY 0.5), so: closed loop over the first 10 steps, then teacher-forced epsilon on the oracle's own trajectory
    assert oracle.psnr_float(tr["x"][10].cpu(), ref["x"][10]) >= PSNR_BAR
    ts = oracle.ddim_timesteps(1000, 250)
    assert len(set(ts.tolist())) == 250
    for i in (11, 60, 125, 200, 249):
        t_b = torch.full((2,), int(ts[i]), dtype=torch.long).cuda()
        e = net(ref["x"][i].cuda(), z, t_b).cpu()
        assert oracle.rel_l2(e, ref["eps"][i]) < EPS_TOL, (i, oracle.rel_l2(e, ref["eps"][i]))


def test_fused_groupnorm_paths_match_unfused(oracle, monkeypatch):
    """CLPK_FUSE_GN=3 (opt-in, see plan.cu): norm2 + SiLU inside conv2 and out_norm inside the `out` conv (row-slab levels,
    W >= 128) against the default plan with the stand-alone GroupNorm passes, and both against the fp32 oracle."""
    cfg = dict(z_dim=512, base=64, ch_mult=(1, 2))
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 128, 128, generator=g)
    z = torch.nn.functional.normalize(torch.randn(2, 512, generator=g), dim=-1)
    t = torch.tensor([999, 250])
    net, sd = make_net(oracle, cfg, seed=5)
    monkeypatch.setenv("CLPK_RES16", "0")                         # these switches act on the plan with the fp32 residual stream
    monkeypatch.setenv("CLPK_FUSE_GN", "3")
    fused = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.setenv("CLPK_FUSE_GN", "0")
    monkeypatch.setenv("CLPK_HEAD16", "0")
    net.release_plans()
    plain = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.delenv("CLPK_FUSE_GN")
    monkeypatch.delenv("CLPK_HEAD16")
    net.release_plans()
    head = net(x.cuda(), z.cuda(), t.cuda()).cpu()                # the default plan: fused head kernel (head_conv.cu)
    assert oracle.rel_l2(head, plain) < 2e-3 and not torch.equal(head, plain)
    monkeypatch.setenv("CLPK_HEAD_FUSED", "0")
    net.release_plans()
    head16 = net(x.cuda(), z.cuda(), t.cuda()).cpu()              # 16-bit-only transposed-conv output, stand-alone out_norm
    monkeypatch.delenv("CLPK_HEAD_FUSED")
    net.release_plans()
    assert oracle.rel_l2(head16, plain) < 2e-3 and oracle.rel_l2(head16, head) < 2e-3 and not torch.equal(head16, head)
    assert not torch.equal(fused, plain)                         # the switch really changes the executed path
    with torch.no_grad():
        ref = oracle.unet_forward(sd, cfg["ch_mult"], x, z, t)
    assert oracle.rel_l2(fused, plain) < 2e-3
    assert oracle.rel_l2(fused, ref) < EPS_TOL and oracle.rel_l2(plain, ref) < EPS_TOL
    assert oracle.rel_l2(head, ref) < EPS_TOL and oracle.rel_l2(head16, ref) < EPS_TOL


def test_16bit_residual_stream_matches_fp32_stream(oracle, monkeypatch):
    """Default plan (fp16 operands): the residual stream of the wide levels (rows >= 128 px) lives in fp16 only (CLPK_RES16,
    plan.cu).  Against the plan with the fp32 stream (CLPK_RES16=0), the plan with the 16-bit stream at every level, and the
    fp32 oracle; bf16 operands keep the fp32 stream (the switch must not change their result)."""
    cfg = dict(z_dim=512, base=64, ch_mult=(1, 2, 2))
    g = torch.Generator().manual_seed(14)
    x = torch.randn(2, 3, 128, 128, generator=g)
    z = torch.nn.functional.normalize(torch.randn(2, 512, generator=g), dim=-1)
    t = torch.tensor([999, 3])
    net, sd = make_net(oracle, cfg, seed=15)
    r16 = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.setenv("CLPK_RES16", "0")
    net.release_plans()
    r32 = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.delenv("CLPK_RES16")
    net.release_plans()
    with torch.no_grad():
        ref = oracle.unet_forward(sd, cfg["ch_mult"], x, z, t)
    monkeypatch.setenv("CLPK_RES16_MIN_W", "0")                     # the 16-bit stream at EVERY level (default: rows >= 128 px)
    net.release_plans()
    rall = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.delenv("CLPK_RES16_MIN_W")
    net.release_plans()
    assert not torch.equal(r16, r32) and not torch.equal(rall, r16)  # the switches really change the executed path
    e16, e32, eall = oracle.rel_l2(r16, ref), oracle.rel_l2(r32, ref), oracle.rel_l2(rall, ref)
    assert e32 < EPS_TOL and e16 < EPS_TOL and eall < EPS_TOL and e16 < 2.0 * e32 + 1e-3 and eall < 2.0 * e32 + 1e-3, (e16, e32, eall)
    assert oracle.rel_l2(r16, r32) < 3e-3 and oracle.rel_l2(rall, r32) < 3e-3
    net.operand_dtype = torch.bfloat16
    b_on = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    monkeypatch.setenv("CLPK_RES16", "0")
    net.release_plans()
    b_off = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    assert torch.equal(b_on, b_off)


def test_ddim_update_fused_into_head_matches_standalone_kernel(oracle, monkeypatch):
    """Default plan: the head kernel applies the DDIM update x <- f(x, eps) in its epilogue (CLPK_HEAD_DDIM, plan.cu).  Same
    element math, same Philox counters as ddim_step_kernel -> the sampled images are bit-identical to the plan that runs
    the stand-alone update kernel, deterministic (eta = 0) and stochastic (in-kernel noise and caller-supplied noise)."""
    cfg = dict(z_dim=512, base=64, ch_mult=(1, 2))
    net, _ = make_net(oracle, cfg, seed=31, out_gain=0.1)
    g = torch.Generator().manual_seed(32)
    z = torch.nn.functional.normalize(torch.randn(2, 512, generator=g), dim=-1).cuda()
    x_T = torch.randn(2, 3, 64, 64, generator=g).cuda()
    noise = torch.randn(6, 2, 3, 64, 64, generator=g).cuda()

    def runs():
        out = [_sampler(0.0).sample(net, z, (2, 3, 64, 64), steps=6, x_T=x_T)]
        s = _sampler(1e-3)
        s.seed = 5
        out.append(s.sample(net, z, (2, 3, 64, 64), steps=6, x_T=x_T))
        out.append(_sampler(1e-3).sample(net, z, (2, 3, 64, 64), steps=6, x_T=x_T, noise=noise))
        tr = {}
        out.append(_sampler(0.0).sample(net, z, (2, 3, 64, 64), steps=6, x_T=x_T, trace=tr))
        out += [tr["eps"], tr["x"]]
        return [o.clone() for o in out]

    fused = runs()
    monkeypatch.setenv("CLPK_HEAD_DDIM", "0")
    net.release_plans()
    plain = runs()
    monkeypatch.delenv("CLPK_HEAD_DDIM")
    net.release_plans()
    for a, b in zip(fused, plain):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    assert not torch.equal(fused[0], fused[1])      # the stochastic path really adds noise


@pytest.mark.parametrize("base,mult,b,h,w", [
    (32, (1, 2), 2, 16, 16),          # tiny image: 8 x 8 at the coarse level, tiles mostly masked
    (32, (1, 2), 3, 40, 40),          # 40-pixel rows: 32-pixel tiles with masked columns at every level
    (32, (1, 2), 2, 72, 56),          # non-square, ragged boxes
    (64, (1, 2), 1, 136, 264),        # row-slab level with a partial second segment (264 = 2 x 128 + 8), 16-bit stream level
    (64, (1, 2, 2), 2, 264, 136),     # three levels, tall image, widths 136 / 68 / 34
])
def test_non_square_and_ragged_image_sizes(oracle, base, mult, b, h, w):
    """Image sizes that are not multiples of the 128-pixel tile (nor square): every level's masking / partial TMA boxes, the
    fused statistics' element counts and the head kernel's bands against the fp32 oracle."""
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    sd = oracle.make_state_dict(512, base, mult, seed=1, out_gain=1.0)
    net = CLIPCondUNet(z_dim=512, base=base, ch_mult=mult)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(b, 3, h, w, generator=g)
    z = torch.nn.functional.normalize(torch.randn(b, 512, generator=g), dim=-1)
    t = torch.randint(0, 1000, (b,), generator=g)
    e = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    with torch.no_grad():
        ref = oracle.unet_forward(sd, mult, x, z, t)
    assert torch.isfinite(e).all()
    for i in range(b):
        assert oracle.rel_l2(e[i], ref[i]) < EPS_TOL / 4, (i, oracle.rel_l2(e[i], ref[i]))


def test_unsupported_base_width_is_a_clean_error(oracle):
    """The plan needs base % 32 == 0 (16-bit K-major rows, 32-column epilogue chunks): anything else must raise, not
    fall back or corrupt memory."""
    from clip_neural_image_conpression_b200._lib import ClpkError
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    net = CLIPCondUNet(z_dim=512, base=48, ch_mult=(1, 2)).cuda().eval()
    with pytest.raises(ClpkError, match="multiple of 32"):
        net(torch.randn(1, 3, 64, 64).cuda(), torch.randn(1, 512).cuda(), torch.tensor([5]).cuda())
