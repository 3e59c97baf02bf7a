"""Host-side logic of the product package on CPU: parameter tree, scheduler tables, DDIM coefficient table, .clp IO,
sharding arithmetic.  (No kernel can run here; everything compared against the golden vectors / the oracle.)"""
import numpy as np
import pytest
import torch

from clip_neural_image_conpression_b200 import parallel
from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
from clip_neural_image_conpression_b200.diffusion.ddim import ddim_coefficients, ddim_timesteps
from clip_neural_image_conpression_b200.io import bitstream
from clip_neural_image_conpression_b200.models import CLIPCondUNet


@pytest.mark.parametrize("cfg", [dict(z_dim=512, base=32, ch_mult=(1, 2)), dict(z_dim=512, base=128, ch_mult=(1, 2, 2)),
                                 dict(z_dim=768, base=64, ch_mult=(1, 2, 2, 4))])
def test_parameter_tree_matches_reference_contract(oracle, cfg):
    net = CLIPCondUNet(**cfg)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == [(k, tuple(s)) for k, s in oracle.param_shapes(cfg["z_dim"], cfg["base"], cfg["ch_mult"])]
    assert net.down_chs[0] == cfg["base"] and len(net.down_chs) == len(cfg["ch_mult"]) + 1
    net.load_state_dict(oracle.make_state_dict(cfg["z_dim"], cfg["base"], cfg["ch_mult"], seed=1), strict=True)


def test_default_architecture_parameter_count():
    assert sum(p.numel() for p in CLIPCondUNet().parameters()) == 32_530_435


def test_scheduler_tables_bit_exact_vs_reference_golden(golden):
    g = golden("scheduler")
    for sch in ("cosine", "linear"):
        s = NoiseScheduler(1000, sch, "cpu")
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas", "posterior_variance"):
            assert np.array_equal(getattr(s, k).numpy(), g[f"{sch}.{k}"]), (sch, k)
    assert s.timesteps == 1000 and s.schedule == "linear" and s.device == "cpu"
    with pytest.raises(ValueError, match="Unknown schedule"):
        NoiseScheduler(10, "sigmoid", "cpu")


def test_scheduler_helpers(oracle):
    s = NoiseScheduler(1000, "cosine", "cpu")
    g = torch.Generator().manual_seed(0)
    x0, n = torch.randn(3, 3, 4, 4, generator=g), torch.randn(3, 3, 4, 4, generator=g)
    t = torch.tensor([0, 500, 999])
    xt = s.q_sample(x0, t, n)
    tabs = oracle.scheduler_tables()
    ref = tabs["sqrt_alphas_cumprod"][t].view(-1, 1, 1, 1) * x0 + tabs["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1) * n
    assert torch.equal(xt, ref)
    assert torch.allclose(s.predict_x0_from_eps(xt, t, n)[:2], x0[:2], atol=1e-4)
    mean, var, x0p = s.p_mean_variance(lambda x, z, tt: n, xt, None, t)
    assert mean.shape == x0.shape and var.shape == (3, 1, 1, 1) and float(x0p.abs().max()) <= 1.0


def test_ddim_coefficient_table_equals_oracle_update(oracle):
    """The host-side table feeds the fused kernel; applying it in plain fp32 must reproduce ddim.py:36-45 exactly."""
    tabs = oracle.scheduler_tables()
    s = NoiseScheduler(1000, "cosine", "cpu")
    g = torch.Generator().manual_seed(0)
    x, e, nz = (torch.randn(64, generator=g) for _ in range(3))
    for steps in (10, 50, 250):
        ts = ddim_timesteps(1000, steps)
        assert torch.equal(ts, oracle.ddim_timesteps(1000, steps)) and int(ts[0]) == 999 and int(ts[-1]) == 0
        for eta in (0.0, 1e-3, 1.0):
            coef = ddim_coefficients(s, ts, eta)
            assert coef.shape == (steps, 5) and coef.dtype == torch.float32
            for i in (0, 1, steps // 2, steps - 1):
                c = coef[i]
                x0 = ((x - c[0] * e) / c[1]).clamp(-1, 1)
                mine = c[2] * x0 + c[3] * e
                if eta > 0 and c[4] > 0:
                    mine = mine + c[4] * nz
                a_t = tabs["alphas_cumprod"][ts[i]]
                a_s = tabs["alphas_cumprod_prev"][ts[i]] if i < steps - 1 else torch.tensor(1.0)
                ref = oracle.ddim_update(x, e, a_t, a_s, eta, nz)
                assert torch.equal(torch.isnan(mine), torch.isnan(ref))
                assert torch.equal(mine[~torch.isnan(ref)], ref[~torch.isnan(ref)]), (steps, eta, i)
    assert torch.isnan(ddim_coefficients(s, ddim_timesteps(1000, 50), 1.0)[0, 3])   # the reference's eta=1 defect


def test_sampler_interface():
    s = DDIMSampler(NoiseScheduler(device="cpu"), eta=0.25)
    assert s.eta == 0.25 and s.sch.timesteps == 1000


def test_bitstream_round_trip_and_reference_files(tmp_path, golden):
    g = golden("quantizer")
    for i in range(4):                       # files written by the reference's writer
        p = tmp_path / f"ref{i}.clp"
        p.write_bytes(g[f"clp{i}"].tobytes())
        assert np.array_equal(bitstream.read_bitstream(p), g["codes"][i])
    for d in (1, 512, 768, 4096):            # our writer -> our reader
        q = np.random.default_rng(d).integers(0, 256, d, dtype=np.uint8)
        bitstream.write_bitstream(q.tobytes(), d, tmp_path / "a.clp")
        blob = (tmp_path / "a.clp").read_bytes()
        assert blob[:4] == b"CLPF" and int.from_bytes(blob[4:8], "little") == len(blob) - 8
        assert np.array_equal(bitstream.read_bitstream(tmp_path / "a.clp"), q)
    # byte-identical container for identical codes (same libzstd, level 22)
    bitstream.write_bitstream(g["codes"][0].tobytes(), 512, tmp_path / "b.clp")
    assert (tmp_path / "b.clp").read_bytes() == g["clp0"].tobytes()
    (tmp_path / "bad.clp").write_bytes(b"NOPE" + blob[4:])
    with pytest.raises(AssertionError, match="Bad magic"):
        bitstream.read_bitstream(tmp_path / "bad.clp")
    paths = []
    for i in range(7):
        bitstream.write_bitstream(g["codes"][i].tobytes(), 512, tmp_path / f"m{i}.clp")
        paths.append(tmp_path / f"m{i}.clp")
    assert np.array_equal(bitstream.read_bitstreams(paths, threads=3), g["codes"][:7])
    assert bitstream.read_bitstreams([]).shape[0] == 0
    # batched writer == the reference's per-vector loop (encode_images.py:79-83): byte-identical files, any thread count
    wpaths = [tmp_path / f"w{i}.clp" for i in range(7)]
    bitstream.write_bitstreams(g["codes"][:7], wpaths, threads=3)
    for a, b in zip(paths, wpaths):
        assert a.read_bytes() == b.read_bytes()
    assert wpaths[0].read_bytes() == g["clp0"].tobytes()            # == the file the reference's own writer produced
    bitstream.write_bitstreams(np.zeros((0, 512), np.uint8), [])
    with pytest.raises(ValueError, match="do not match"):
        bitstream.write_bitstreams(g["codes"][:3], wpaths[:2])


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [parallel.shard_bounds(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 4, 4)


def test_reference_arm_runs_the_staged_reference(tmp_path):
    """bench.py's CPU arm: with oracle/_ref present it times the UNMODIFIED reference (zipimport of the archive written by
    oracle/stage_ref.py); the archive holds exactly the reference's .py files.  64 px keeps this at a few seconds."""
    import importlib.util
    import zipfile
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    archive = root / "oracle" / "_ref" / "clip_feature_codec.zip"
    if not archive.exists():
        if not Path("/root/reference/src/clip_feature_codec").exists():
            pytest.skip("neither the staged archive nor the reference tree is present")
        from oracle import stage_ref
        assert stage_ref.stage()
    names = zipfile.ZipFile(archive).namelist()
    assert "clip_feature_codec/models/unet.py" in names and "clip_feature_codec/diffusion/ddim.py" in names
    ref_tree = Path("/root/reference/src/clip_feature_codec")
    if ref_tree.exists():      # byte-identical to the sources where they lie
        z = zipfile.ZipFile(archive)
        for n in names:
            assert z.read(n) == (ref_tree / n.split("/", 1)[1]).read_bytes(), n
    spec = importlib.util.spec_from_file_location("bench_mod", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ips, ms, cores, kind = bench.cpu_reference_sample(64, 1, 1, 0)
    assert kind == "reference" and ips > 0 and ms > 0 and cores >= 1


def test_bench_extras_are_declared():
    """The bench line's contract keys for the extra configurations exist in the source (cheap guard against renames)."""
    from pathlib import Path
    src = (Path(__file__).resolve().parents[1] / "bench.py").read_text()
    for key in ("bf16_operands", "config3_store1024", "config4_ddim250_b64", "config5_wide_512px", "cpu_baseline", "roofline",
                "h2d_bytes_per_step", "gpu_launches"):
        assert key in src
