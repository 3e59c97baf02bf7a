"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (imported from /root/reference/src).

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/gen_golden.py
The fixtures pin oracle/codec_oracle.py (tests/test_oracle_golden.py) and are the reference-side truth of the GPU
parity tests.  All seeds live here; the reference has none.  `zstandard` (pyproject.toml:21) is not installed in this
image, so a shim module over the system libzstd is injected before PKG/io/bitstream.py is imported.
"""
from __future__ import annotations

import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF_SRC = Path("/root/reference/src")
sys.path.insert(0, str(ROOT))

from oracle import codec_oracle as O  # noqa: E402


def install_zstandard_shim() -> None:
    if "zstandard" in sys.modules:
        return
    m = types.ModuleType("zstandard")

    class ZstdCompressor:
        def __init__(self, level=3):
            self.level = level

        def compress(self, data):
            return O.zstd_compress(bytes(data), self.level)

    class ZstdDecompressor:
        def decompress(self, data):
            return O.zstd_decompress(bytes(data))

    m.ZstdCompressor, m.ZstdDecompressor = ZstdCompressor, ZstdDecompressor
    sys.modules["zstandard"] = m


def import_reference():
    if not REF_SRC.exists():
        raise SystemExit("reference not present at /root/reference — golden vectors can only be generated in the build container")
    install_zstandard_shim()
    sys.path.insert(0, str(REF_SRC))
    import clip_feature_codec.codecs.quantizer as rq
    import clip_feature_codec.diffusion.ddim as rd
    import clip_feature_codec.diffusion.scheduler as rs
    import clip_feature_codec.eval.metrics as rm
    import clip_feature_codec.io.bitstream as rb
    import clip_feature_codec.models.blocks as rbl
    import clip_feature_codec.models.unet as ru
    return types.SimpleNamespace(q=rq, ddim=rd, sched=rs, metrics=rm, bits=rb, blocks=rbl, unet=ru)


# configurations of the fixtures (small enough to commit; the GPU tests re-run the same seeds)
TINY = dict(z_dim=512, base=32, ch_mult=(1, 2))        # BASELINE.json configs[0] architecture
MID = dict(z_dim=512, base=64, ch_mult=(1, 2))         # the reference's own unit-test architecture (tests/test_unet.py:8)


def ref_net(R, cfg, seed, out_gain=1.0):
    net = R.unet.CLIPCondUNet(z_dim=cfg["z_dim"], base=cfg["base"], ch_mult=cfg["ch_mult"])
    net.load_state_dict(O.make_state_dict(cfg["z_dim"], cfg["base"], cfg["ch_mult"], seed=seed, out_gain=out_gain), strict=True)
    return net.eval()


def write_ddpm(R, out) -> None:
    """DDPM helpers of the reference scheduler (PKG/diffusion/scheduler.py:46-68) on CPU -> tests/golden/ddpm.npz."""
    g = torch.Generator().manual_seed(13)
    x0 = torch.randn(4, 3, 16, 16, generator=g) * 0.8
    noise, eps = torch.randn(4, 3, 16, 16, generator=g), torch.randn(4, 3, 16, 16, generator=g)
    t = torch.tensor([0, 17, 500, 999])
    dd = {"x0": x0.numpy(), "noise": noise.numpy(), "eps": eps.numpy(), "t": t.numpy()}
    for sch in ("cosine", "linear"):
        s = R.sched.NoiseScheduler(1000, sch, "cpu")
        xt = s.q_sample(x0, t, noise)
        mean, var, x0c = s.p_mean_variance(lambda x, z, tt: eps, xt, None, t)
        dd.update({f"{sch}.xt": xt.numpy(), f"{sch}.x0_pred": s.predict_x0_from_eps(xt, t, eps).numpy(),
                   f"{sch}.mean": mean.numpy(), f"{sch}.var": var.numpy(), f"{sch}.x0_clamped": x0c.numpy()})
    np.savez_compressed(out / "ddpm.npz", **dd)


DEFAULT = dict(z_dim=512, base=128, ch_mult=(1, 2, 2))          # BASELINE.json configs[1-3] architecture
WIDE = dict(z_dim=768, base=192, ch_mult=(1, 2, 2, 4))          # BASELINE.json configs[4] architecture
FULL_KEEP = (0, 1, 2, 25, 49)                                    # DDIM-50 steps recorded: t = 999, 978, 958, 489, 0


def write_fullsize(R, out) -> None:
    """Reference-CPU fixtures at the REAL sizes (tests/golden/fullsize_default.npz, fullsize_wide.npz):
      * default UNet (base 128, (1,2,2)), 256 px, B = 1: the unmodified reference DDIMSampler.sample (ddim.py:21-46)
        over all 50 steps with contractive `out.*` (x0.1); the model callable records (x_t, eps) at steps FULL_KEEP
        (teacher-forced per-step parity at t = 999 / 978 / 958 / 489 / 0) and the final x (closed-loop >= 40 dB bar).
        x_T and z are regenerated from their seeds by the tests, not stored.
      * wide UNet (base 192, (1,2,2,4), z = 768), 128 px, B = 1: two forwards (t = 999, 123) of unet.py:81-106.
    Weights: oracle.make_state_dict(seed) loaded into the reference module (strict)."""
    torch.set_num_threads(8)
    sch = R.sched.NoiseScheduler(1000, "cosine", "cpu")
    net = ref_net(R, DEFAULT, seed=0, out_gain=0.1)
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(1, 512, generator=g), dim=-1)
    x_T = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    rec, calls = {}, [0]

    def model(x, zc, t):
        i = calls[0]
        calls[0] += 1
        eps = net(x, zc, t)
        if i in FULL_KEEP:
            if i > 0:
                rec[f"x{i}"] = x.detach().clone().numpy()
            rec[f"eps{i}"] = eps.detach().clone().numpy()
            rec[f"t{i}"] = t.numpy().copy()
        return eps

    xf = R.ddim.DDIMSampler(sch, 0.0).sample(model, z, (1, 3, 256, 256), steps=50, x_T=x_T)
    assert calls[0] == 50
    rec["x_final"] = xf.numpy()
    rec["z_check"], rec["x_T_check"] = z.numpy(), x_T.numpy()[:, :, :2, :8]   # guards the seeds, not the payload
    np.savez_compressed(out / "fullsize_default.npz", **rec)
    del net

    net = ref_net(R, WIDE, seed=31, out_gain=0.1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, 128, 128, generator=g)
    zw = torch.nn.functional.normalize(torch.randn(1, 768, generator=g), dim=-1)
    wd = {}
    with torch.no_grad():
        for t in (999, 123):
            wd[f"eps_t{t}"] = net(x, zw, torch.tensor([t])).numpy()
    wd["z_check"], wd["x_check"] = zw.numpy(), x.numpy()[:, :, :2, :8]
    np.savez_compressed(out / "fullsize_wide.npz", **wd)
    for f in ("fullsize_default.npz", "fullsize_wide.npz"):
        print(f, (out / f).stat().st_size // 1024, "KiB")


def main() -> None:
    R = import_reference()
    if "--only-fullsize" in sys.argv:   # adds the full-size fixtures without rewriting the others
        write_fullsize(R, ROOT / "tests" / "golden")
        return
    if "--only-ddpm" in sys.argv:   # adds the ddpm fixture without rewriting the others
        write_ddpm(R, ROOT / "tests" / "golden")
        return
    out = ROOT / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)

    # ---- scheduler tables (PKG/diffusion/scheduler.py)
    tabs = {}
    for sch in ("cosine", "linear"):
        s = R.sched.NoiseScheduler(1000, sch, "cpu")
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas", "posterior_variance"):
            tabs[f"{sch}.{k}"] = getattr(s, k).numpy()
    np.savez_compressed(out / "scheduler.npz", **tabs)

    # ---- timestep embedding (PKG/models/unet.py:22-39)
    t = torch.tensor([0, 1, 20, 489, 500, 978, 999])
    np.savez_compressed(out / "timestep_embedding.npz", t=t.numpy(), emb256=R.unet.timestep_embedding(t, 256).numpy(),
                        emb64=R.unet.timestep_embedding(t, 64).numpy())

    # ---- quantiser + bitstream (PKG/codecs/quantizer.py, PKG/io/bitstream.py, reconstruct_diffusion.py:39-44)
    g = torch.Generator().manual_seed(5)
    Z = torch.nn.functional.normalize(torch.randn(24, 512, generator=g), dim=-1)
    Z[3] = Z[3] * 0.0  # an all-zero row: exercises the max(norm, 1e-9) guard after dequantisation of zero-ish data
    quant = R.q.PerChannelAffineQuantizer(8).fit(Z)
    codes = np.stack([quant.encode(Z[i]) for i in range(Z.shape[0])])
    dec = quant.decode(codes)
    sys.path.insert(0, str(REF_SRC))
    from clip_feature_codec.cli.reconstruct_diffusion import l2_normalize_np  # needs PIL only
    scale, zero = quant.scale.numpy().astype("float32"), quant.zero.numpy().astype("float32")
    z_dec = np.stack([l2_normalize_np((codes[i].astype(np.float32) * scale + zero)[None, :]).astype(np.float32)[0]
                      for i in range(codes.shape[0])])
    blobs = []
    with tempfile.TemporaryDirectory() as td:
        for i in range(4):
            p = Path(td) / f"{i}.clp"
            R.bits.write_bitstream(codes[i].tobytes(), 512, p)
            assert (R.bits.read_bitstream(p) == codes[i]).all()
            blobs.append(np.frombuffer(p.read_bytes(), dtype=np.uint8))
    np.savez_compressed(out / "quantizer.npz", Z=Z.numpy(), scale=scale, zero=zero, codes=codes, decoded=dec, z_dec=z_dec,
                        **{f"clp{i}": b for i, b in enumerate(blobs)})

    # ---- UNet forward, FiLM, ResBlock (PKG/models/unet.py:81-106, blocks.py)
    fw = {}
    for name, cfg, size in (("tiny", TINY, 32), ("mid", MID, 64)):
        net = ref_net(R, cfg, seed=11)
        g = torch.Generator().manual_seed(12)
        x = torch.randn(2, 3, size, size, generator=g)
        z = torch.nn.functional.normalize(torch.randn(2, cfg["z_dim"], generator=g), dim=-1)
        tt = torch.tensor([999, 37])
        with torch.no_grad():
            fw[f"{name}.eps"] = net(x, z, tt).numpy()
        fw[f"{name}.x"], fw[f"{name}.z"], fw[f"{name}.t"] = x.numpy(), z.numpy(), tt.numpy()
    torch.manual_seed(3)
    film = R.blocks.FiLM(16, 32)
    rbk = R.blocks.ResBlock(32, 256)
    g = torch.Generator().manual_seed(4)
    xf, hf = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 32, generator=g)
    xr, hr = torch.randn(2, 32, 16, 16, generator=g), torch.randn(2, 256, generator=g)
    with torch.no_grad():
        fw["film.y"], fw["rb.y"] = film(xf, hf).numpy(), rbk(xr, hr).numpy()
    fw["film.x"], fw["film.h"], fw["rb.x"], fw["rb.h"] = xf.numpy(), hf.numpy(), xr.numpy(), hr.numpy()
    for k, v in film.state_dict().items():
        fw["film.sd." + k] = v.numpy()
    for k, v in rbk.state_dict().items():
        fw["rb.sd." + k] = v.numpy()
    np.savez_compressed(out / "unet_forward.npz", **fw)

    # ---- DDIM (PKG/diffusion/ddim.py): BASELINE configs[0] = tiny arch, 64 px, batch 2, 10 steps
    dd = {}
    sch = R.sched.NoiseScheduler(1000, "cosine", "cpu")
    g = torch.Generator().manual_seed(5)
    z = torch.nn.functional.normalize(torch.randn(2, 512, generator=g), dim=-1)
    x_T = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(6))
    dd["z"], dd["x_T"] = z.numpy(), x_T.numpy()
    for tag, gain in (("p0", 1.0), ("pg", 0.1)):
        net = ref_net(R, TINY, seed=0, out_gain=gain)
        dd[f"{tag}.eta0.x"] = R.ddim.DDIMSampler(sch, 0.0).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T).numpy()
    net = ref_net(R, TINY, seed=0, out_gain=0.1)
    # eta = 1e-3: finite stochastic path; the reference draws torch.randn_like from the global generator
    torch.manual_seed(7)
    dd["pg.eta1e-3.x"] = R.ddim.DDIMSampler(sch, 1e-3).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T).numpy()
    torch.manual_seed(7)
    dd["noise_seed7"] = torch.stack([torch.randn(2, 3, 64, 64) for _ in range(10)]).numpy()[:, :1, :1, :4, :4]  # spot check
    # eta = 1.0: the reference returns all-NaN
    torch.manual_seed(7)
    x_nan = R.ddim.DDIMSampler(sch, 1.0).sample(net, z, (2, 3, 64, 64), steps=10, x_T=x_T)
    dd["pg.eta1.nan_fraction"] = np.array(float(torch.isnan(x_nan).float().mean()))
    # 50-step contractive run (parity bar: >= 40 dB vs this output)
    dd["pg.eta0.steps50.x"] = R.ddim.DDIMSampler(sch, 0.0).sample(net, z, (2, 3, 64, 64), steps=50, x_T=x_T).numpy()
    np.savez_compressed(out / "ddim.npz", **dd)

    # ---- output post-process + PSNR (reconstruct_diffusion.py:55-56, PKG/eval/metrics.py:16-29)
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(3, 3, 32, 32, generator=g) * 0.7).numpy()
    b = (torch.from_numpy(a) + 0.05 * torch.randn(3, 3, 32, 32, generator=g)).numpy().astype(np.float32)
    u8 = np.stack([((np.clip(a[i], -1, 1).transpose(1, 2, 0) + 1.0) * 127.5).astype(np.uint8) for i in range(3)])
    ps = np.array([R.metrics.psnr(a[i], b[i]) for i in range(3)] + [R.metrics.psnr(a[0], a[0])])
    np.savez_compressed(out / "metrics.npz", a=a, b=b, u8=u8, psnr=ps, metric_u8=R.metrics._to_uint8(a))
    write_ddpm(R, out)
    total = sum(p.stat().st_size for p in out.glob("*.npz"))
    print(f"wrote {len(list(out.glob('*.npz')))} fixtures, {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
