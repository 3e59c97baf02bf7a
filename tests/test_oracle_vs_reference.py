"""Live comparison oracle <-> unmodified reference (only where /root/reference exists, i.e. the build container)
plus: the product's nn.Module tree initialises exactly like the reference's under the same seed."""
import numpy as np
import torch


def test_seeded_init_matches_reference(reference):
    from clip_neural_image_conpression_b200.models import CLIPCondUNet
    torch.manual_seed(0)
    ref = reference.unet.CLIPCondUNet(z_dim=512, base=32, ch_mult=(1, 2))
    torch.manual_seed(0)
    mine = CLIPCondUNet(z_dim=512, base=32, ch_mult=(1, 2))
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert ref.down_chs == mine.down_chs


def test_oracle_unet_equals_reference_on_fresh_inputs(reference, oracle):
    torch.manual_seed(1)
    ref = reference.unet.CLIPCondUNet(z_dim=64, base=32, ch_mult=(2, 1)).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    x, z, t = torch.randn(3, 3, 16, 16), torch.randn(3, 64), torch.tensor([0, 400, 999])
    with torch.no_grad():
        assert oracle.rel_l2(oracle.unet_forward(sd, (2, 1), x, z, t), ref(x, z, t)) < 1e-6


def test_oracle_ddim_equals_reference(reference, oracle):
    torch.manual_seed(2)
    ref = reference.unet.CLIPCondUNet(z_dim=64, base=32, ch_mult=(1,)).eval()
    with torch.no_grad():
        ref.out.weight.mul_(0.1)
        ref.out.bias.mul_(0.1)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    sch = reference.sched.NoiseScheduler(1000, "linear", "cpu")
    tabs = oracle.scheduler_tables(1000, "linear")
    z, x_T = torch.randn(2, 64), torch.randn(2, 3, 16, 16)
    for eta in (0.0, 5e-4):
        torch.manual_seed(3)
        a = reference.ddim.DDIMSampler(sch, eta).sample(ref, z, (2, 3, 16, 16), steps=7, x_T=x_T)
        torch.manual_seed(3)
        noise = torch.stack([torch.randn(2, 3, 16, 16) for _ in range(7)])
        with torch.no_grad():
            b = oracle.ddim_sample(lambda x, zc, t: oracle.unet_forward(sd, (1,), x, zc, t), tabs, z, x_T, 7, eta, noise)
        assert oracle.psnr_float(a, b) > 100.0


def test_reference_reader_accepts_our_files(reference, tmp_path):
    from clip_neural_image_conpression_b200.io import bitstream
    q = np.random.default_rng(0).integers(0, 256, 768, dtype=np.uint8)
    bitstream.write_bitstream(q.tobytes(), 768, tmp_path / "x.clp")
    assert np.array_equal(reference.bits.read_bitstream(tmp_path / "x.clp"), q)
    reference.bits.write_bitstream(q.tobytes(), 768, tmp_path / "y.clp")
    assert np.array_equal(bitstream.read_bitstream(tmp_path / "y.clp"), q)
