"""End-to-end on the GPU through the reference-facing surface: synthetic store on disk (.clp + codec_meta.npz +
manifest.json + weights .pt) -> reconstruct_diffusion / eval CLIs."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_store(tmp_path, oracle, n=5, size=32, base=32, ch_mult=(1, 2), orig_hw=None):
    from clip_neural_image_conpression_b200.codecs import PerChannelAffineQuantizer
    from clip_neural_image_conpression_b200.io.bitstream import write_bitstream
    from PIL import Image
    g = torch.Generator().manual_seed(5)
    Z = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1)
    qz = PerChannelAffineQuantizer(8).fit(Z)
    np.savez(tmp_path / "codec_meta.npz", scale=qz.scale.cpu().numpy(), zero=qz.zero.cpu().numpy(), dim=np.int32(512))
    manifest = []
    rng = np.random.default_rng(0)
    for i in range(n):
        hw = orig_hw[i % len(orig_hw)] if orig_hw else (size, size)   # originals of other sizes exercise the BICUBIC resize
        img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
        Image.fromarray(img).save(tmp_path / f"{i}.png")
        write_bitstream(qz.encode(Z[i]).tobytes(), 512, tmp_path / f"{i}.clp")
        manifest.append({"image": str(tmp_path / f"{i}.png"), "bitstream": str(tmp_path / f"{i}.clp")})
    (tmp_path / "manifest.json").write_text(json.dumps(manifest))
    sd = oracle.make_state_dict(512, base, ch_mult, seed=0, out_gain=0.1)
    torch.save(sd, tmp_path / "w.pt")
    return Z, qz, sd, manifest


def test_cli_reconstruct_and_eval(tmp_path, oracle, capsys):
    from clip_neural_image_conpression_b200.cli import eval as cli_eval
    from clip_neural_image_conpression_b200.cli import reconstruct_diffusion as cli_rec
    from PIL import Image
    Z, qz, sd, manifest = make_store(tmp_path, oracle)
    common = ["--store_dir", str(tmp_path), "--weights", str(tmp_path / "w.pt"), "--size", "32", "--steps", "5",
              "--base", "32", "--ch_mult", "1", "2", "--seed", "0"]
    cli_rec.main(common + ["--bitstream", manifest[0]["bitstream"], "--out", str(tmp_path / "r.png")])
    assert "Saved to" in capsys.readouterr().out
    img = np.array(Image.open(tmp_path / "r.png"))
    assert img.shape == (32, 32, 3) and img.dtype == np.uint8
    # same seed -> the oracle reproduces the image from the same x_T (CUDA generator) within 1 grey level
    torch.manual_seed(0)
    x_T = torch.randn((1, 3, 32, 32), device="cuda").cpu()
    codes = qz.encode(Z[0])
    z = torch.from_numpy(oracle.l2_normalize(oracle.dequant(codes, qz.scale.cpu().numpy(), qz.zero.cpu().numpy())[None]))
    tabs = oracle.scheduler_tables(1000, "cosine")
    with torch.no_grad():
        ref = oracle.ddim_sample(lambda x, zc, t: oracle.unet_forward(sd, (1, 2), x, zc, t), tabs, z, x_T, steps=5)
    ref_u8 = oracle.to_uint8_image(ref[0].numpy())
    assert np.abs(img.astype(int) - ref_u8.astype(int)).max() <= 2
    assert (img == ref_u8).mean() > 0.9

    cli_eval.main(common + ["--batch", "2", "--out_json", str(tmp_path / "m.json")])   # 5 images, ragged last batch
    out = capsys.readouterr().out
    assert "Average PSNR:" in out and "Average SSIM:" in out and "Average LPIPS:" in out and "Average CLIP similarity:" in out
    rows = json.loads((tmp_path / "m.json").read_text())
    assert len(rows) == 5 and set(rows[0]) == {"image", "psnr", "ssim", "lpips", "clip_sim"}
    assert "Average SSIM: nan" not in out
    # VALUE parity with the reference's loop body (eval.py:56-79) restated by the oracle: same codes, same x_T draws
    # (--seed 0 -> the CUDA generator yields one [2,3,32,32] draw per micro-batch, the padded slot of the last one unused)
    ref_rows = oracle_eval_rows(oracle, Z, qz, sd, manifest, seed=0, batch=2, size=32, steps=5)
    for r, (pr, sr) in zip(rows, ref_rows):
        assert abs(r["psnr"] - pr) < 0.02, (r["psnr"], pr)          # dB; ours differs by fp16 operands (<= 2 grey levels)
        assert abs(r["ssim"] - sr) < 2e-3, (r["ssim"], sr)
    avg_p = float(np.mean([p for p, _ in ref_rows]))
    avg_s = float(np.mean([s_ for _, s_ in ref_rows]))
    got_p = float(out.split("Average PSNR:")[1].split("dB")[0])
    got_s = float(out.split("Average SSIM:")[1].split()[0])
    assert abs(got_p - avg_p) < 0.02 and abs(got_s - avg_s) < 2e-3


def oracle_eval_rows(oracle, Z, qz, sd, manifest, seed, batch, size, steps, rank_offset=0, lo=0, hi=None):
    """(psnr, ssim) per image of manifest[lo:hi] computed by the oracle's restatement of eval.py:56-79, with the x_T the
    CLI draws for that shard (torch.manual_seed(seed + rank), one CUDA draw per micro-batch)."""
    from clip_neural_image_conpression_b200.cli.eval import load_original
    hi = len(manifest) if hi is None else hi
    torch.manual_seed(seed + rank_offset)
    n = hi - lo
    x_T = torch.cat([torch.randn((batch, 3, size, size), device="cuda") for _ in range((n + batch - 1) // batch)]).cpu()
    tabs = oracle.scheduler_tables(1000, "cosine")
    scale, zero = qz.scale.cpu().numpy(), qz.zero.cpu().numpy()
    rows = []
    for j, i in enumerate(range(lo, hi)):
        q = oracle.clp_decode(open(manifest[i]["bitstream"], "rb").read())
        z = torch.from_numpy(oracle.l2_normalize(oracle.dequant(q, scale, zero)[None]))
        with torch.no_grad():
            x = oracle.ddim_sample(lambda xx, zc, t: oracle.unet_forward(sd, (1, 2), xx, zc, t), tabs, z, x_T[j:j + 1], steps=steps)
        rec = x[0].clamp(-1, 1).numpy()
        img0 = load_original(manifest[i]["image"], size)
        rows.append((oracle.psnr(img0, rec), oracle.ssim(img0, rec)))
    return rows


def test_cli_eval_resizes_originals_like_pillow(tmp_path, oracle, capsys):
    """Originals larger / smaller / non-square than --size: the CLI's device-side BICUBIC resize gives the metric values of
    the reference loop body, whose originals go through Pillow's resize on the host (eval.py:66-67)."""
    from clip_neural_image_conpression_b200.cli import eval as cli_eval
    Z, qz, sd, manifest = make_store(tmp_path, oracle, n=4, orig_hw=[(48, 40), (20, 33), (32, 32), (100, 64)])
    cli_eval.main(["--store_dir", str(tmp_path), "--weights", str(tmp_path / "w.pt"), "--size", "32", "--steps", "5", "--base", "32",
                   "--ch_mult", "1", "2", "--seed", "0", "--batch", "4", "--out_json", str(tmp_path / "m.json")])
    capsys.readouterr()
    rows = json.loads((tmp_path / "m.json").read_text())
    ref_rows = oracle_eval_rows(oracle, Z, qz, sd, manifest, seed=0, batch=4, size=32, steps=5)
    for r, (pr, sr) in zip(rows, ref_rows):
        assert abs(r["psnr"] - pr) < 0.02 and abs(r["ssim"] - sr) < 2e-3, (r, pr, sr)


def test_cli_eval_two_ranks_nccl(tmp_path, oracle):
    """The torchrun / NCCL branch of the eval CLI on 2 GPUs: contiguous manifest shards, all-reduced metric sums,
    all_gather_object of the rows — values equal the oracle's per-shard loop.  Skipped on single-GPU boxes."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    Z, qz, sd, manifest = make_store(tmp_path, oracle)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", "-m", "clip_neural_image_conpression_b200.cli.eval", "--store_dir", str(tmp_path),
           "--weights", str(tmp_path / "w.pt"), "--size", "32", "--steps", "5", "--base", "32", "--ch_mult", "1", "2",
           "--seed", "0", "--batch", "2", "--out_json", str(tmp_path / "m2.json")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = json.loads((tmp_path / "m2.json").read_text())
    assert [row["image"] for row in rows] == [m["image"] for m in manifest]      # gathered in manifest order
    ref = oracle_eval_rows(oracle, Z, qz, sd, manifest, 0, 2, 32, 5, rank_offset=0, lo=0, hi=3) + \
        oracle_eval_rows(oracle, Z, qz, sd, manifest, 0, 2, 32, 5, rank_offset=1, lo=3, hi=5)
    for row, (pr, sr) in zip(rows, ref):
        assert abs(row["psnr"] - pr) < 0.02 and abs(row["ssim"] - sr) < 2e-3
    got_p = float(r.stdout.split("Average PSNR:")[1].split("dB")[0])
    assert abs(got_p - float(np.mean([p for p, _ in ref]))) < 0.02


def test_decode_codes_matches_per_image_decoding(tmp_path, oracle):
    """Batched + padded micro-batches give the same images as decoding each code alone (the reference's B = 1 loop)."""
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    from clip_neural_image_conpression_b200.io.bitstream import read_bitstreams
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    from clip_neural_image_conpression_b200.pipeline import decode_codes
    Z, qz, sd, manifest = make_store(tmp_path, oracle)
    net = CLIPCondUNet(512, 32, (1, 2))
    net.load_state_dict(sd)
    net = net.cuda().eval()
    sampler = DDIMSampler(NoiseScheduler(1000, "cosine", "cuda"), 0.0)
    q = read_bitstreams([m["bitstream"] for m in manifest])
    assert q.shape == (5, 512) and q.dtype == np.uint8
    x_T = torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(6)).cuda()
    scale, zero = qz.scale.cuda(), qz.zero.cuda()
    full = decode_codes(net, sampler, q, scale, zero, 32, steps=5, batch=4, x_T=x_T)
    solo = torch.cat([decode_codes(net, sampler, q[i:i + 1], scale, zero, 32, steps=5, batch=1, x_T=x_T[i:i + 1]) for i in range(5)])
    assert oracle.psnr_float(full.cpu(), solo.cpu()) > 60.0
    assert decode_codes(net, sampler, q[:0], scale, zero, 32, steps=5, batch=4).shape == (0, 3, 32, 32)


def test_write_store_matches_the_reference_store_layout(tmp_path, oracle, golden):
    """pipeline.write_store = the store-writing tail of the reference's encode CLI (encode_images.py:75-87) with the
    quantiser on the device: codec_meta.npz, every .clp file and manifest.json equal what the oracle's restatement of the
    reference arithmetic + the reference's container format produce — bit for bit — and the store decodes back."""
    from clip_neural_image_conpression_b200.io.bitstream import read_bitstreams
    from clip_neural_image_conpression_b200.pipeline import write_store
    g = torch.Generator().manual_seed(9)
    n, d = 37, 512
    feats = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).numpy()
    feats[3, 7] = feats[:, 7].max() + 0.5                            # an outlier sets one channel's range
    imgs = [str(tmp_path / "imgs" / f"pic_{i:03d}.jpg") for i in range(n)]
    manifest = write_store(feats, imgs, tmp_path / "store", threads=5)
    scale, zero = oracle.quant_fit(feats)
    meta = np.load(tmp_path / "store" / "codec_meta.npz")
    assert set(meta.files) == {"scale", "zero", "dim"} and meta["dim"].dtype == np.int32 and int(meta["dim"]) == d
    assert meta["scale"].dtype == np.float32 and np.array_equal(meta["scale"], scale) and np.array_equal(meta["zero"], zero)
    on_disk = json.loads((tmp_path / "store" / "manifest.json").read_text(encoding="utf-8"))
    assert on_disk == manifest and len(manifest) == n
    assert manifest[5] == {"image": imgs[5], "bitstream": str(tmp_path / "store" / "pic_005.clp")}
    codes = read_bitstreams([m["bitstream"] for m in manifest])
    ref_codes = np.stack([oracle.quant_encode(feats[i], scale, zero) for i in range(n)])
    assert np.array_equal(codes, ref_codes)
    for i in (0, 17, 36):                                            # container bytes == the oracle's restatement of the format
        assert Path(manifest[i]["bitstream"]).read_bytes() == oracle.clp_encode(ref_codes[i].tobytes())
    with pytest.raises(SystemExit):
        write_store(np.zeros((0, d), np.float32), [], tmp_path / "empty")


def test_cli_encode_features_then_eval_round_trip(tmp_path, oracle, capsys):
    """encode_features CLI (store writer) -> eval CLI (store reader + decoder): the codec round trip through files only."""
    from clip_neural_image_conpression_b200.cli import encode_features as cli_enc
    from clip_neural_image_conpression_b200.cli import eval as cli_eval
    from PIL import Image
    g = torch.Generator().manual_seed(10)
    n = 3
    feats = torch.nn.functional.normalize(torch.randn(n, 512, generator=g), dim=-1).numpy()
    np.save(tmp_path / "f.npy", feats)
    imgs = []
    rng = np.random.default_rng(1)
    for i in range(n):
        Image.fromarray(rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)).save(tmp_path / f"im{i}.png")
        imgs.append(str(tmp_path / f"im{i}.png"))
    (tmp_path / "list.txt").write_text("\n".join(imgs) + "\n")
    cli_enc.main(["--features", str(tmp_path / "f.npy"), "--image_list", str(tmp_path / "list.txt"), "--out_dir", str(tmp_path / "store")])
    assert f"Done. Stored {n} vectors in" in capsys.readouterr().out
    torch.save(oracle.make_state_dict(512, 32, (1, 2), seed=0, out_gain=0.1), tmp_path / "w.pt")
    cli_eval.main(["--store_dir", str(tmp_path / "store"), "--weights", str(tmp_path / "w.pt"), "--size", "32", "--steps", "3",
                   "--base", "32", "--ch_mult", "1", "2", "--seed", "0", "--batch", "2", "--out_json", str(tmp_path / "m.json")])
    rows = json.loads((tmp_path / "m.json").read_text())
    assert [r["image"] for r in rows] == imgs and all(np.isfinite(r["psnr"]) for r in rows)
