"""The oracle (oracle/codec_oracle.py) against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, generator: oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

TINY = dict(z_dim=512, base=32, ch_mult=(1, 2))
MID = dict(z_dim=512, base=64, ch_mult=(1, 2))


def test_scheduler_tables_bit_exact(oracle, golden):
    g = golden("scheduler")
    for sch in ("cosine", "linear"):
        tabs = oracle.scheduler_tables(1000, sch)
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas", "posterior_variance"):
            assert np.array_equal(tabs[k].numpy(), g[f"{sch}.{k}"]), (sch, k)


def test_timestep_embedding_bit_exact(oracle, golden):
    g = golden("timestep_embedding")
    t = torch.from_numpy(g["t"])
    assert np.array_equal(oracle.timestep_embedding(t, 256).numpy(), g["emb256"])
    assert np.array_equal(oracle.timestep_embedding(t, 64).numpy(), g["emb64"])


def test_quantizer_bit_exact(oracle, golden):
    g = golden("quantizer")
    scale, zero = oracle.quant_fit(g["Z"])
    assert np.array_equal(scale, g["scale"]) and np.array_equal(zero, g["zero"])
    codes = oracle.quant_encode(g["Z"], scale, zero)
    assert np.array_equal(codes, g["codes"])
    assert np.array_equal(oracle.dequant(codes, scale, zero), g["decoded"])
    assert np.array_equal(oracle.l2_normalize(oracle.dequant(codes, scale, zero)), g["z_dec"])


def test_clp_container(oracle, golden):
    g = golden("quantizer")
    for i in range(4):
        blob = g[f"clp{i}"].tobytes()
        assert blob[:4] == b"CLPF"
        assert np.array_equal(oracle.clp_decode(blob), g["codes"][i])
        # our writer -> same container; frames from the same libzstd are byte identical
        assert oracle.clp_encode(g["codes"][i].tobytes()) == blob
    try:
        oracle.clp_decode(b"XXXX" + blob[4:])
        raise RuntimeError("bad magic accepted")
    except AssertionError as e:
        assert "Bad magic" in str(e)


def test_unet_forward_matches_reference(oracle, golden):
    g = golden("unet_forward")
    for name, cfg in (("tiny", TINY), ("mid", MID)):
        sd = oracle.make_state_dict(cfg["z_dim"], cfg["base"], cfg["ch_mult"], seed=11)
        with torch.no_grad():
            eps = oracle.unet_forward(sd, cfg["ch_mult"], torch.from_numpy(g[f"{name}.x"]), torch.from_numpy(g[f"{name}.z"]),
                                      torch.from_numpy(g[f"{name}.t"]))
        # same ATen kernels, same op order -> identical up to thread-partitioning of reductions
        assert oracle.rel_l2(eps, torch.from_numpy(g[f"{name}.eps"])) < 1e-6


def test_resblock_and_film_match_reference(oracle, golden):
    g = golden("unet_forward")
    sd = {"b." + k[len("rb.sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("rb.sd.")}
    with torch.no_grad():
        y = oracle._resblock(sd, "b", torch.from_numpy(g["rb.x"]), torch.from_numpy(g["rb.h"]))
    assert oracle.rel_l2(y, torch.from_numpy(g["rb.y"])) < 1e-6


def test_ddim_matches_reference(oracle, golden):
    g = golden("ddim")
    tabs = oracle.scheduler_tables(1000, "cosine")
    z, x_T = torch.from_numpy(g["z"]), torch.from_numpy(g["x_T"])
    for tag, gain in (("p0", 1.0), ("pg", 0.1)):
        sd = oracle.make_state_dict(512, 32, (1, 2), seed=0, out_gain=gain)
        fn = lambda x, zc, t: oracle.unet_forward(sd, (1, 2), x, zc, t)  # noqa: E731
        with torch.no_grad():
            x = oracle.ddim_sample(fn, tabs, z, x_T, steps=10, eta=0.0)
        assert oracle.psnr_float(x, torch.from_numpy(g[f"{tag}.eta0.x"])) > 100.0, tag
    sd = oracle.make_state_dict(512, 32, (1, 2), seed=0, out_gain=0.1)
    fn = lambda x, zc, t: oracle.unet_forward(sd, (1, 2), x, zc, t)  # noqa: E731
    torch.manual_seed(7)
    noise = torch.stack([torch.randn(2, 3, 64, 64) for _ in range(10)])
    assert np.array_equal(noise.numpy()[:, :1, :1, :4, :4], g["noise_seed7"])
    with torch.no_grad():
        x = oracle.ddim_sample(fn, tabs, z, x_T, steps=10, eta=1e-3, noise=noise)
        assert oracle.psnr_float(x, torch.from_numpy(g["pg.eta1e-3.x"])) > 100.0
        x = oracle.ddim_sample(fn, tabs, z, x_T, steps=10, eta=1.0, noise=noise)
    assert float(torch.isnan(x).float().mean()) == float(g["pg.eta1.nan_fraction"]) == 1.0


def test_metrics_match_reference(oracle, golden):
    g = golden("metrics")
    for i in range(3):
        assert np.array_equal(oracle.to_uint8_image(g["a"][i]), g["u8"][i])
        assert oracle.psnr(g["a"][i], g["b"][i]) == g["psnr"][i]
    assert oracle.psnr(g["a"][0], g["a"][0]) == float("inf") == g["psnr"][3]
    assert np.array_equal(oracle.metric_uint8(g["a"]), g["metric_u8"])


def test_ssim_known_answers(oracle):
    """SSIM (metrics.py:32-46) is restated from scikit-image's published algorithm (parity unpinned: skimage is not
    vendored); anchor it on closed-form answers and structural properties."""
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (3, 33, 47)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.2, a.shape), -1, 1).astype(np.float32)
    assert oracle.ssim(a, a) == 1.0                                           # identical images
    assert oracle.ssim(a, b) == oracle.ssim(b, a) and 0.0 < oracle.ssim(a, b) < 1.0
    # constant images: zero variances -> S = (2 ux uy + C1) / (ux^2 + uy^2 + C1) everywhere
    ca, cb = np.full((3, 16, 16), -0.2, np.float32), np.full((3, 16, 16), 0.5, np.float32)
    ua, ub = float(oracle.metric_uint8(ca)[0, 0, 0]), float(oracle.metric_uint8(cb)[0, 0, 0])
    c1 = (0.01 * 255) ** 2
    np.testing.assert_allclose(oracle.ssim(ca, cb), (2 * ua * ub + c1) / (ua * ua + ub * ub + c1), rtol=1e-12)
    # brute-force 7x7 windows (no filter library) on a small plane, interior pixels only
    x, y = oracle.metric_uint8(a[0]).astype(np.float64), oracle.metric_uint8(b[0]).astype(np.float64)
    vals = []
    for i in range(3, x.shape[0] - 3):
        for j in range(3, x.shape[1] - 3):
            wx, wy = x[i - 3:i + 4, j - 3:j + 4], y[i - 3:i + 4, j - 3:j + 4]
            ux, uy = wx.mean(), wy.mean()
            vx, vy = wx.var(ddof=1), wy.var(ddof=1)
            vxy = ((wx - ux) * (wy - uy)).sum() / 48.0
            vals.append(((2 * ux * uy + c1) * (2 * vxy + (0.03 * 255) ** 2)) /
                        ((ux * ux + uy * uy + c1) * (vx + vy + (0.03 * 255) ** 2)))
    np.testing.assert_allclose(oracle.ssim(a[:1].repeat(3, 0), b[:1].repeat(3, 0)), np.mean(vals), rtol=1e-10)
    with pytest.raises(ValueError):
        oracle.ssim(a[:, :5], b[:, :5])                                       # smaller than the 7x7 window


def test_ddpm_helpers_match_reference(oracle, golden):
    """q_sample / predict_x0_from_eps / p_mean_variance (scheduler.py:46-68): oracle == reference output, bit for bit."""
    g = golden("ddpm")
    x0, noise, eps, t = (torch.from_numpy(g[k]) for k in ("x0", "noise", "eps", "t"))
    for sch in ("cosine", "linear"):
        tabs = oracle.scheduler_tables(1000, sch)
        xt = oracle.q_sample(tabs, x0, t, noise)
        assert np.array_equal(xt.numpy(), g[f"{sch}.xt"])
        assert np.array_equal(oracle.predict_x0_from_eps(tabs, xt, t, eps).numpy(), g[f"{sch}.x0_pred"])
        mean, var, x0c = oracle.p_mean_variance(tabs, eps, xt, t)
        assert np.array_equal(mean.numpy(), g[f"{sch}.mean"]) and np.array_equal(var.numpy(), g[f"{sch}.var"])
        assert np.array_equal(x0c.numpy(), g[f"{sch}.x0_clamped"])


@pytest.mark.parametrize("h,w,oh,ow", [(37, 53, 32, 32), (300, 200, 256, 256), (64, 64, 256, 256), (256, 256, 256, 256),
                                        (17, 9, 64, 48), (500, 30, 31, 300), (256, 300, 256, 128)])
def test_bicubic_restatement_matches_pillow(oracle, h, w, oh, ow):
    """The oracle's restatement of Pillow's 8-bit BICUBIC resampler (third-party arithmetic behind PKG/cli/eval.py:66) and
    the product's host-side coefficient tables, against Pillow itself: every byte / every coefficient identical."""
    from PIL import Image
    from clip_neural_image_conpression_b200.eval.resample import bicubic_coeffs
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(oracle.bicubic_resize_u8(img, oh, ow), ref)
    for in_size, out_size in ((w, ow), (h, oh)):
        b, k = bicubic_coeffs(in_size, out_size)
        ob, ok = oracle._pil_bicubic_coeffs(in_size, out_size)
        assert np.array_equal(b, np.asarray(ob, np.int32)) and np.array_equal(k, np.asarray(ok, np.int32))
    f = oracle.original_to_float_chw(ref)
    assert f.dtype == np.float32 and f.shape == (3, oh, ow) and np.array_equal(f, (ref.astype(np.float32) / 127.5 - 1.0).transpose(2, 0, 1))
