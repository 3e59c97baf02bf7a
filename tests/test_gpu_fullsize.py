"""GPU parity at the REAL sizes and at hostile value ranges.

* Full-size fixtures written by the unmodified reference on the CPU (oracle/gen_golden.py --only-fullsize):
  default UNet (base 128, ch_mult (1,2,2)) at 256 px, B = 1 — teacher-forced epsilon at DDIM-50 steps 0/1/2/25/49
  (t = 999/978/958/489/0) <= 1e-2 relative L2, and the closed 50-step loop >= 40 dB against the reference's own fp32
  result; wide UNet (base 192, ch_mult (1,2,2,4), z 768) at 128 px — two forwards <= 1e-2.
* Range tests: fp16 copies of un-normalised tensors saturate instead of overflowing to inf; GroupNorm statistics stay
  accurate for |mean| >> std (shifted sums, fp64 combination) in both the stand-alone kernel and the conv epilogue."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEFAULT = dict(z_dim=512, base=128, ch_mult=(1, 2, 2))
WIDE = dict(z_dim=768, base=192, ch_mult=(1, 2, 2, 4))
EPS_TOL = 1e-2
PSNR_BAR = 40.0
KEEP = (0, 1, 2, 25, 49)


def make_net(oracle, cfg, seed, out_gain=1.0):
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    net = CLIPCondUNet(z_dim=cfg["z_dim"], base=cfg["base"], ch_mult=cfg["ch_mult"])
    sd = oracle.make_state_dict(cfg["z_dim"], cfg["base"], cfg["ch_mult"], seed=seed, out_gain=out_gain)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), sd


@pytest.mark.parametrize("operand", [torch.float16, torch.bfloat16])
def test_default_256px_vs_reference_fixture(oracle, golden, operand):
    """BASELINE configs[1] network and image size against outputs of the reference itself (CPU fp32)."""
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    g = golden("fullsize_default")
    net, _ = make_net(oracle, DEFAULT, seed=0, out_gain=0.1)
    net.operand_dtype = operand
    z = F.normalize(torch.randn(1, 512, generator=torch.Generator().manual_seed(5)), dim=-1)
    x_T = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    assert np.array_equal(z.numpy(), g["z_check"]) and np.array_equal(x_T.numpy()[:, :, :2, :8], g["x_T_check"])
    ts = oracle.ddim_timesteps(1000, 50)
    worst = 0.0
    for i in KEEP:                                                   # teacher forced on the reference's own x_t
        x = x_T if i == 0 else torch.from_numpy(g[f"x{i}"])
        assert int(g[f"t{i}"][0]) == int(ts[i])
        eps = net(x.cuda(), z.cuda(), torch.full((1,), int(ts[i]), dtype=torch.long, device="cuda")).cpu()
        err = oracle.rel_l2(eps, torch.from_numpy(g[f"eps{i}"]))
        worst = max(worst, err)
        assert err < EPS_TOL, (i, err)
    if operand == torch.float16:
        assert worst < 3e-3, worst                                   # fp16 operands sit well inside the bar
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", "cuda"), eta=0.0)
    x = sampler.sample(net, z.cuda(), (1, 3, 256, 256), steps=50, x_T=x_T.cuda()).cpu()
    psnr = oracle.psnr_float(x, torch.from_numpy(g["x_final"]))
    assert psnr >= PSNR_BAR, psnr                                    # closed loop, contractive weights
    # the same image inside a batch of 8 (the benchmarked plan): row 3 of the batch == the B = 1 result
    zb = F.normalize(torch.randn(8, 512, generator=torch.Generator().manual_seed(7)), dim=-1)
    xb = torch.randn(8, 3, 256, 256, generator=torch.Generator().manual_seed(8))
    zb[3], xb[3] = z[0], x_T[0]
    x8 = sampler.sample(net, zb.cuda(), (8, 3, 256, 256), steps=50, x_T=xb.cuda()).cpu()
    assert oracle.psnr_float(x8[3:4], torch.from_numpy(g["x_final"])) >= PSNR_BAR
    net.release_plans()
    torch.cuda.empty_cache()


def test_wide_128px_vs_reference_fixture(oracle, golden):
    """BASELINE configs[4] architecture at 128 px (W = 128 -> row-slab mainloop, 24/48-channel GroupNorm groups summed in
    pieces, 384-channel rows, 3072-channel bottom level) against the reference's CPU output."""
    g = golden("fullsize_wide")
    net, _ = make_net(oracle, WIDE, seed=31, out_gain=0.1)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, 128, 128, generator=gen)
    z = F.normalize(torch.randn(1, 768, generator=gen), dim=-1)
    assert np.array_equal(z.numpy(), g["z_check"]) and np.array_equal(x.numpy()[:, :, :2, :8], g["x_check"])
    for t in (999, 123):
        eps = net(x.cuda(), z.cuda(), torch.tensor([t], device="cuda")).cpu()
        err = oracle.rel_l2(eps, torch.from_numpy(g[f"eps_t{t}"]))
        assert err < EPS_TOL, (t, err)
    net.release_plans()
    del net
    torch.cuda.empty_cache()


def test_config4_batch64_plan_matches_batch1(oracle, golden):
    """BASELINE configs[3] plan shape (default UNet, 256 px, batch 64): the B = 64 plan (other tile counts, GroupNorm slot
    counts and grid sizes than B = 1 / 8) reproduces the reference's teacher-forced epsilon for the fixture image
    placed at three batch positions, and the stochastic graph loop (eta = 1e-3, in-graph Philox noise) stays finite."""
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    g = golden("fullsize_default")
    net, _ = make_net(oracle, DEFAULT, seed=0, out_gain=0.1)
    z1 = F.normalize(torch.randn(1, 512, generator=torch.Generator().manual_seed(5)), dim=-1)
    zb = F.normalize(torch.randn(64, 512, generator=torch.Generator().manual_seed(9)), dim=-1)
    xb = torch.randn(64, 3, 256, 256, generator=torch.Generator().manual_seed(10))
    x25, ref = torch.from_numpy(g["x25"]), torch.from_numpy(g["eps25"])
    for pos in (0, 37, 63):
        zb[pos], xb[pos] = z1[0], x25[0]
    eps = net(xb.cuda(), zb.cuda(), torch.full((64,), 489, dtype=torch.long, device="cuda")).cpu()
    for pos in (0, 37, 63):
        assert oracle.rel_l2(eps[pos:pos + 1], ref) < EPS_TOL
    assert torch.isfinite(eps).all()
    s = DDIMSampler(NoiseScheduler(1000, "cosine", "cuda"), eta=1e-3)
    s.seed = 3
    x = s.sample(net, zb.cuda(), (64, 3, 256, 256), steps=5, x_T=xb.cuda())
    assert torch.isfinite(x).all()
    net.release_plans()
    del net
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ value ranges
def test_fp16_stores_saturate_instead_of_overflowing():
    """A conv output beyond the fp16 range is stored as +-65504 in the 16-bit copy (cvt.rn.satfinite), never as inf."""
    from clip_neural_image_conpression_b200 import ops
    g = torch.Generator().manual_seed(1)
    xb = torch.randn(1, 16, 16, 64, generator=g).to(torch.float16).cuda()
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda()
    bias = torch.full((64,), 1.0e5).cuda()
    bias[::2] = -1.0e5
    bias[5] = 0.25
    o = ops.conv_igemm(xb, ops.pack_conv_weight(wt, 0), 0, 64, bias, want_f32=True, want_op=True)
    assert torch.isfinite(o["op"]).all()
    assert float(o["f32"].abs().max()) > 9e4                       # the fp32 stream keeps the true value
    ref = o["f32"].clamp(-65504.0, 65504.0).to(torch.float16)
    assert torch.equal(o["op"], ref)
    only = ops.conv_igemm(xb, ops.pack_conv_weight(wt, 0), 0, 64, bias, want_f32=False, want_op=True)   # staged 16-bit path
    assert torch.equal(only["op"], ref)


@pytest.mark.parametrize("what", ["residual", "film"])
def test_unet_beyond_fp16_range_stays_finite_and_accurate(oracle, what):
    """Un-normalised tensors pushed past the fp16 maximum: "residual" scales the stem until the residual stream (and its
    16-bit copy X16 that feeds the stride-2 / transposed convs) peaks at 8e4; "film" scales the FiLM of the first blocks
    until the conv1 + FiLM output Y peaks at 8e4.  Without saturating stores the outliers become inf and the next
    GroupNorm turns the image into NaN.  GroupNorm is scale invariant, so the fp32 oracle's output barely moves, and ours
    must stay inside the epsilon bar (the saturated outliers are a ~1e-5 fraction of the elements)."""
    cfg = dict(z_dim=512, base=64, ch_mult=(1, 2))
    net, sd = make_net(oracle, cfg, seed=11)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 3, 64, 64, generator=g)
    z = F.normalize(torch.randn(2, 512, generator=g), dim=-1)
    t = torch.tensor([999, 37])
    sd = {k: v.clone() for k, v in sd.items()}
    taps = {}
    with torch.no_grad():
        oracle.unet_forward(sd, cfg["ch_mult"], x, z, t, taps=taps)
    if what == "residual":
        stem = F.conv2d(x, sd["in_conv.weight"], sd["in_conv.bias"], padding=1)
        gain = 8.0e4 / float(stem.abs().max())
        sd["in_conv.weight"] *= gain
        sd["in_conv.bias"] *= gain
    else:
        for blk in ("down.0", "down.1", "down.3"):
            gain = 8.0e4 / float(taps[blk + ".y"].abs().max())       # Y' = gain * Y:  1 + s' = gain (1 + s), b' = gain b
            sd[blk + ".film.to_scale.weight"] *= gain
            sd[blk + ".film.to_scale.bias"] = sd[blk + ".film.to_scale.bias"] * gain + (gain - 1.0)
            sd[blk + ".film.to_shift.weight"] *= gain
            sd[blk + ".film.to_shift.bias"] *= gain
    net.load_state_dict(sd, strict=True)
    eps = net(x.cuda(), z.cuda(), t.cuda()).cpu()
    assert torch.isfinite(eps).all()
    with torch.no_grad():
        taps = {}
        ref = oracle.unet_forward(sd, cfg["ch_mult"], x, z, t, taps=taps)
    probe = taps["pre_out"] if what == "residual" else taps["down.0.y"]
    assert float(probe.abs().max()) > 7e4                            # the reference's tensor really is that large
    for b in range(2):
        assert oracle.rel_l2(eps[b], ref[b]) < EPS_TOL, (what, b, oracle.rel_l2(eps[b], ref[b]))


@pytest.mark.parametrize("mean,std", [(500.0, 1.0), (-3.0e3, 2.0), (0.0, 1.0)])
def test_groupnorm_large_mean(mean, std):
    """GroupNorm statistics for |mean| / std = 500 ... 1500 against F.group_norm in fp64 (torch uses Welford; a plain
    one-pass E[x^2] - E[x]^2 in fp32 loses the variance entirely here)."""
    from clip_neural_image_conpression_b200 import ops
    g = torch.Generator().manual_seed(2)
    b, h, w, c = 2, 48, 40, 64
    x = (mean + std * torch.randn(b, h, w, c, generator=g)).cuda()
    x[..., 8:16] *= -1.0                                            # one group with the opposite sign
    gamma, beta = (1 + 0.1 * torch.randn(c, generator=g)).cuda(), (0.1 * torch.randn(c, generator=g)).cuda()
    y = ops.groupnorm_silu(x, gamma, beta, 8, 1e-5, True).float()
    ref = F.silu(F.group_norm(x.permute(0, 3, 1, 2).double(), 8, gamma.double(), beta.double(), 1e-5)).permute(0, 2, 3, 1)
    assert float((y.double() - ref).abs().max()) < 3e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("kind,b,h,w,cin,cout", [
    (0, 2, 32, 32, 128, 128),      # generic mainloop, 16 channels per group
    (0, 1, 5, 256, 64, 64),        # row-slab mainloop, 8 channels per group
    (0, 1, 24, 24, 64, 64),        # ragged tiles (masked rows)
    (2, 1, 16, 16, 128, 64),       # transposed conv (4 phases)
    (0, 1, 16, 16, 512, 512),      # 64 channels per group: two 32-column chunks per group
])
def test_conv_fused_statistics_large_mean(kind, b, h, w, cin, cout):
    """Fused conv-epilogue statistics with the output mean 500x its standard deviation == fp64 statistics of the conv's own
    fp32 output (rstd to 1e-3: the fp32 output itself only resolves ~3e-5 around 500)."""
    from clip_neural_image_conpression_b200 import ops
    g = torch.Generator().manual_seed(3 + cout)
    xb = torch.randn(b, h, w, cin, generator=g).to(torch.float16).cuda()
    wt = ((torch.randn(cin, cout, 4, 4, generator=g) if kind == 2 else torch.randn(cout, cin, 3, 3, generator=g))
          / (cin * (4 if kind == 2 else 9)) ** 0.5).cuda()
    bias = torch.full((cout,), 500.0).cuda()
    bias[cout // 2:] = -500.0
    o = ops.conv_igemm(xb, ops.pack_conv_weight(wt, kind), kind, cout, bias, gn_groups=8)
    oh, ow = o["f32"].shape[1:3]
    y = o["f32"].double().reshape(b, oh * ow, 8, cout // 8)
    mean = y.mean(dim=(1, 3))
    rstd = 1.0 / torch.sqrt(y.var(dim=(1, 3), unbiased=False) + 1e-5)
    np.testing.assert_allclose(o["gn_stats"][..., 0].cpu().numpy(), mean.cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(o["gn_stats"][..., 1].cpu().numpy(), rstd.cpu().numpy(), rtol=1e-3)
