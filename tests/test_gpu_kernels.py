"""GPU parity of every leaf entry point of libclpk.so (called through the C ABI) against the oracle and the golden
vectors of the unmodified reference.  Integer / byte work: bit exact.  Floating point: tolerance stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from clip_neural_image_conpression_b200 import ops as o
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------ codec (bit exact)
def test_dequant_bit_exact_and_l2norm(ops, oracle, golden):
    g = golden("quantizer")
    z, raw = ops.dequant_l2norm(cu(g["codes"]), cu(g["scale"]), cu(g["zero"]), l2norm=True, return_raw=True)
    assert np.array_equal(raw.cpu().numpy(), g["decoded"])                     # bit exact (north-star)
    # L2 normalisation: numpy's pairwise fp32 sum vs a warp-tree fp32 sum -> the norm may differ in the last ulp
    np.testing.assert_allclose(z.cpu().numpy(), g["z_dec"], rtol=3e-7, atol=1e-9)
    z2 = ops.dequant_l2norm(cu(g["codes"]), cu(g["scale"]), cu(g["zero"]), l2norm=False)
    assert np.array_equal(z2.cpu().numpy(), g["decoded"])


def test_dequant_edge_cases(ops, oracle):
    rng = np.random.default_rng(0)
    for b, d in ((1, 1), (3, 7), (5, 768), (2, 4099)):
        q = rng.integers(0, 256, (b, d), dtype=np.uint8)
        q[0, :] = 0
        scale = rng.random(d, dtype=np.float32) * 1e-2 + 1e-8
        zero = (rng.standard_normal(d) * 0.1).astype(np.float32)
        if b > 1:
            zero_row = np.zeros(d, np.float32)  # a row that dequantises to exactly 0 -> max(norm, 1e-9) guard
            raw = ops.dequant_l2norm(cu(q[:1]), cu(scale), cu(zero_row), l2norm=True)
            assert torch.all(raw == 0)
        raw = ops.dequant_l2norm(cu(q), cu(scale), cu(zero), l2norm=False).cpu().numpy()
        assert np.array_equal(raw, oracle.dequant(q, scale, zero))
    empty = ops.dequant_l2norm(torch.zeros((0, 16), dtype=torch.uint8, device="cuda"), cu(np.ones(16, np.float32)),
                               cu(np.zeros(16, np.float32)))
    assert empty.shape == (0, 16)


def test_quantizer_fit_encode_bit_exact(ops, golden):
    g = golden("quantizer")
    scale, zero = ops.quant_fit(cu(g["Z"]))
    assert np.array_equal(scale.cpu().numpy(), g["scale"]) and np.array_equal(zero.cpu().numpy(), g["zero"])
    q = ops.quant_encode(cu(g["Z"]), scale, zero)
    assert np.array_equal(q.cpu().numpy(), g["codes"])


def test_quantizer_class_round_trip(golden):
    from clip_neural_image_conpression_b200.codecs import PerChannelAffineQuantizer
    g = golden("quantizer")
    qz = PerChannelAffineQuantizer(8).fit(torch.from_numpy(g["Z"]))
    codes = qz.encode(torch.from_numpy(g["Z"]))
    assert codes.dtype == np.uint8 and np.array_equal(codes, g["codes"])
    assert np.array_equal(qz.decode(codes), g["decoded"])
    # tie cases: values exactly half way between two codes must round half to even like torch.round
    qz.scale, qz.zero = torch.full((4,), 2.0), torch.zeros(4)
    assert qz.encode(torch.tensor([1.0, 3.0, 5.0, 600.0])).tolist() == [0, 2, 2, 255]


# ------------------------------------------------------------------------------------------------ DDIM update
def test_ddim_step_bit_exact(ops, oracle):
    tabs = oracle.scheduler_tables(1000, "cosine")
    g = torch.Generator().manual_seed(0)
    x, eps, nz = (torch.randn(2, 3, 33, 31, generator=g) for _ in range(3))  # odd size: exercises the scalar tail
    from clip_neural_image_conpression_b200.diffusion.ddim import ddim_coefficients, ddim_timesteps
    from clip_neural_image_conpression_b200.diffusion import NoiseScheduler
    sch = NoiseScheduler(1000, "cosine", "cpu")
    for eta in (0.0, 1e-3, 1.0):
        ts = ddim_timesteps(1000, 10)
        coef = ddim_coefficients(sch, ts, eta)
        for i in (0, 1, 5, 9):
            a_t = tabs["alphas_cumprod"][ts[i]]
            a_s = tabs["alphas_cumprod_prev"][ts[i]] if i < 9 else torch.tensor(1.0)
            ref = oracle.ddim_update(x, eps, a_t, a_s, eta, nz)
            got = ops.ddim_step(x.cuda(), eps.cuda(), coef[i].tolist(), nz.cuda() if eta > 0 else None).cpu()
            assert torch.equal(torch.isnan(got), torch.isnan(ref)), (eta, i)      # NaN pattern (eta=1: all NaN early)
            m = ~torch.isnan(ref)
            assert torch.equal(got[m], ref[m]), (eta, i)                           # bit exact elsewhere


# ------------------------------------------------------------------------------------------------ small fp kernels
def test_timestep_embedding(ops, golden):
    g = golden("timestep_embedding")
    for dim, key in ((256, "emb256"), (64, "emb64")):
        e = ops.timestep_embedding(cu(g["t"]), dim).cpu().numpy()
        # frequencies from the torch-CPU table (what the reference evaluates): only sinf / cosf of CUDA vs the CPU's libm
        # differ, by an ulp or two of a value in [-1, 1]
        np.testing.assert_allclose(e, g[key], rtol=0, atol=5e-7)
        # the device's own expf (the reference with t on CUDA): a 1-ulp frequency difference is amplified by t <= 999
        d = ops.timestep_embedding(cu(g["t"]), dim, host_freqs=False).cpu().numpy()
        np.testing.assert_allclose(d, g[key], rtol=0, atol=2e-4)


def test_linear(ops):
    g = torch.Generator().manual_seed(1)
    for m, n, k, act in ((1, 256, 512, 1), (8, 1024, 256, 1), (50, 256, 1024, 0), (13, 7680, 256, 0)):
        x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** 0.5, torch.randn(n, generator=g)
        ref = F.linear(x.double(), w.double(), b.double())
        ref = F.silu(ref) if act else ref
        y = ops.linear(x.cuda(), w.cuda(), b.cuda(), act=act).cpu()
        np.testing.assert_allclose(y.numpy(), ref.float().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_groupnorm_silu(ops, dtype):
    g = torch.Generator().manual_seed(2)
    tol = 1.5e-3 if dtype == torch.float16 else 6e-3   # half an ulp of the output format + fast-exp noise
    for b, h, w, c, silu in ((2, 16, 16, 32, True), (3, 32, 32, 128, True), (1, 8, 8, 512, False), (2, 8, 8, 192, True),
                             (1, 4, 4, 3072, True), (2, 64, 64, 64, True)):
        x = (torch.randn(b, h, w, c, generator=g) * 1.7 + 0.4)
        gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
        ref = F.group_norm(x.permute(0, 3, 1, 2).double(), 8, gamma.double(), beta.double(), 1e-5)
        ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
        y = ops.groupnorm_silu(x.cuda(), gamma.cuda(), beta.cuda(), 8, 1e-5, silu, dtype=dtype).cpu()
        assert y.dtype == dtype
        err = (y.double() - ref).abs()
        assert float((err / (ref.abs() + 1e-2)).max()) < tol, (b, h, w, c)
        y2 = ops.groupnorm_silu(x.cuda(), gamma.cuda(), beta.cuda(), 8, 1e-5, silu, dtype=dtype).cpu()
        assert torch.equal(y, y2)  # deterministic (no float atomics)


def test_conv_in(ops):
    g = torch.Generator().manual_seed(3)
    for b, h, w, cout in ((2, 40, 24, 128), (1, 16, 16, 32), (1, 8, 8, 192)):
        x, wt, bias = torch.randn(b, 3, h, w, generator=g), torch.randn(cout, 3, 3, 3, generator=g) / 5, torch.randn(cout, generator=g)
        ref = F.conv2d(x.double(), wt.double(), bias.double(), padding=1).permute(0, 2, 3, 1).float()
        y = ops.conv_in(x.cuda(), wt.cuda(), bias.cuda()).cpu()
        np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


def test_conv_16bit_only_output_with_fused_statistics(ops):
    """conv1 of a ResBlock as the plan runs it: FiLM epilogue, output kept ONLY in the 16-bit operand format, GroupNorm
    statistics taken from the fp32 accumulators."""
    _conv1_like(ops, 2, 32, 32, 128)
    _conv1_like(ops, 1, 3, 256, 128)     # row-slab kernel: staged 16-bit TMA store from per-warp 64-byte-row slots


def _conv1_like(ops, b, h, w, c):
    g = torch.Generator().manual_seed(3)
    xb = torch.randn(b, h, w, c, generator=g).to(torch.float16).cuda()
    wt = (torch.randn(c, c, 3, 3, generator=g) * 0.03).cuda()
    bias = torch.randn(c, generator=g).cuda()
    sc, sh = (1 + 0.3 * torch.randn(b, c, generator=g)).cuda(), torch.randn(b, c, generator=g).cuda()
    wp = ops.pack_conv_weight(wt, 0)
    full = ops.conv_igemm(xb, wp, 0, c, bias, film_scale1p=sc, film_shift=sh, want_f32=True, want_op=True, gn_groups=8)
    only = ops.conv_igemm(xb, wp, 0, c, bias, film_scale1p=sc, film_shift=sh, want_f32=False, want_op=True, gn_groups=8)
    assert "f32" not in only and torch.equal(only["op"], full["op"]) and torch.equal(only["gn_stats"], full["gn_stats"])
    assert torch.equal(only["op"], full["f32"].to(torch.float16))          # round-to-nearest-even of the fp32 result


def test_stem_im2col_and_pointwise_conv(ops):
    """The stem (unet.py:55) as the plan runs it: im2col (27 -> 32 columns, 16-bit) + CLPK_CONV_1X1 on the tensor cores."""
    import ctypes as C
    from clip_neural_image_conpression_b200 import _lib
    g = torch.Generator().manual_seed(4)
    b, h, w, cout = 2, 24, 40, 128
    x = torch.randn(b, 3, h, w, generator=g).cuda()
    wt = (torch.randn(cout, 3, 3, 3, generator=g) / 5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    cols = torch.empty((b, h, w, 32), dtype=torch.float16, device="cuda")
    _lib.check(_lib.load().clpk_stem_im2col(x.data_ptr(), cols.data_ptr(), b, 3, h, w, _lib.OP_F16, _lib.stream_ptr()), "im2col")
    ref_cols = F.unfold(x, 3, padding=1).reshape(b, 27, h, w).permute(0, 2, 3, 1)          # k = c*9 + r*3 + s
    assert torch.equal(cols[..., :27], ref_cols.to(torch.float16)) and torch.all(cols[..., 27:] == 0)
    wpad = torch.zeros(cout, 32, 1, 1, device="cuda")
    wpad[:, :27, 0, 0] = wt.reshape(cout, 27)
    o = ops.conv_igemm(cols, ops.pack_conv_weight(wpad, _lib.CONV_1X1), _lib.CONV_1X1, cout, bias, gn_groups=8)
    ref = F.conv2d(x.to(torch.float16).double(), wt.to(torch.float16).double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert float((o["f32"].double() - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    exact = F.conv2d(x.double(), wt.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert float(torch.linalg.norm(o["f32"].double() - exact) / torch.linalg.norm(exact)) < 1e-3   # fp16 operand rounding


# ------------------------------------------------------------------------------------------------ tensor-core convs
CONV_CASES = [
    # kind, B, H, W, Cin, Cout, film, resid
    (0, 1, 8, 16, 64, 64, False, False),       # one M tile
    (0, 2, 32, 32, 128, 128, True, False),     # conv1 + FiLM epilogue
    (0, 2, 32, 32, 128, 128, False, True),     # conv2 + residual epilogue
    (0, 1, 16, 16, 32, 32, False, False),      # BLOCK_K = 32 (64-byte swizzle)
    (0, 1, 8, 8, 512, 512, False, True),       # two N tiles, 72 k-blocks, half-filled M tile
    (0, 1, 4, 256, 128, 128, False, False),    # W > 128: row-slab mainloop (shifted smem descriptors), CTA pair
    (0, 3, 5, 192, 64, 64, True, True),        # slab: partial second row tile, odd tile count (masked pair half), 1 k-block
    (0, 1, 3, 128, 192, 128, False, True),     # slab: three channel blocks, residual look-ahead across tiles
    (0, 2, 6, 320, 128, 3, False, False),      # slab, single CTA, narrow N (`out` conv at full width), ragged W
    (0, 1, 2, 256, 64, 192, True, False),      # slab declined (N = 192 > 128): generic CTA-pair path on a wide row
    (0, 1, 9, 40, 64, 64, False, True),        # width that is neither a multiple nor a divisor of 32: 32-pixel tiles
    (0, 1, 24, 24, 64, 64, False, False),      # ragged boxes
    (0, 2, 16, 16, 128, 3, False, False),      # `out` conv: N padded to 16, NCHW output
    (0, 4, 64, 64, 256, 256, True, False),     # several tiles per CTA: smem ring wrap + TMEM double buffering
    (0, 1, 16, 16, 192, 192, False, True),     # wide config channel count (N = 192)
    (1, 1, 32, 32, 64, 128, False, False),     # stride 2
    (1, 2, 16, 16, 128, 256, False, False),
    (1, 1, 16, 16, 32, 64, False, False),
    (2, 1, 8, 8, 128, 64, False, True),        # transposed conv + skip add
    (2, 2, 16, 16, 256, 128, False, True),
    (2, 1, 8, 8, 64, 32, False, False),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,b,h,w,cin,cout,film,resid", CONV_CASES)
def test_conv_igemm(ops, kind, b, h, w, cin, cout, film, resid, dtype):
    g = torch.Generator().manual_seed(kind * 1000 + cin + cout + h)
    xb = torch.randn(b, h, w, cin, generator=g).to(dtype).cuda()
    if kind == 2:
        wt = (torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** 0.5).cuda()
    else:
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    wp = ops.pack_conv_weight(wt, kind, dtype)
    xr, wr = xb.float().permute(0, 3, 1, 2).double(), wt.to(dtype).double()
    if kind == 0:
        ref = F.conv2d(xr, wr, bias.double(), padding=1)
    elif kind == 1:
        ref = F.conv2d(xr, wr, bias.double(), stride=2, padding=1)
    else:
        ref = F.conv_transpose2d(xr, wr, bias.double(), stride=2, padding=1)
    kw = {}
    if film:
        sc, sh = (1 + 0.3 * torch.randn(b, cout, generator=g)).cuda(), torch.randn(b, cout, generator=g).cuda()
        kw.update(film_scale1p=sc, film_shift=sh)
        ref = ref * sc.double()[:, :, None, None] + sh.double()[:, :, None, None]
    if resid:
        r = torch.randn(b, ref.shape[2], ref.shape[3], cout, generator=g).cuda()
        kw["resid"] = r
        ref = ref + r.permute(0, 3, 1, 2).double()
    nchw = cout % 16 != 0
    o = ops.conv_igemm(xb, wp, kind, cout, bias, want_f32=not nchw, want_op=not nchw, want_nchw=nchw, **kw)
    d = ops.conv_direct(xb, wp, kind, cout, bias, want_f32=not nchw, want_op=False, want_nchw=nchw, **kw)
    y = o["nchw"] if nchw else o["f32"].permute(0, 3, 1, 2)
    yd = d["nchw"] if nchw else d["f32"].permute(0, 3, 1, 2)
    scale = float(ref.abs().max())
    # identical 16-bit operands, fp32 accumulation: only the summation order differs from the fp64 reference
    assert float((y.double() - ref).abs().max()) < 1e-4 * max(scale, 1.0)
    assert float((y - yd).abs().max()) < 1e-4 * max(scale, 1.0)
    if not nchw:
        assert o["op"].dtype == dtype
        half_ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
        assert float((o["op"].double().permute(0, 3, 1, 2) - ref).abs().max()) < 1.2 * half_ulp * max(scale, 1.0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("kind,b,h,w,cin,cout,film", [
    (0, 2, 32, 32, 128, 128, False),     # conv2 on the 16-bit residual stream, generic tiles
    (0, 3, 5, 256, 128, 128, True),      # row-slab CTA pairs, odd tile count, several residual sub-boxes in flight
    (0, 1, 8, 8, 512, 512, False),       # two N tiles, half-filled M tile
    (0, 4, 64, 64, 256, 256, False),     # several tiles per CTA pair: slot ring wraps many times
    (0, 1, 9, 40, 64, 64, False),        # 32-pixel tiles, masked columns
    (2, 2, 16, 16, 256, 128, False),     # transposed conv + 16-bit skip add, 4 phases
    (2, 1, 8, 8, 128, 64, False),
    (0, 1, 16, 16, 64, 48, False),       # cout % 32 != 0: direct (un-chunked) epilogue
])
def test_conv_16bit_residual_stream(ops, kind, b, h, w, cin, cout, film, dtype):
    """clpk_conv_epilogue.resid_op: the residual arrives as a 16-bit NHWC tile and the 16-bit sum is stored, in place
    (blocks.py:44 / unet.py:104 on a residual stream kept in 16 bits).  Result == round(fp32 conv + bias + residual) to
    half an ulp of the 16-bit format, fused GroupNorm statistics == those of the UN-rounded sums, and the cross-check
    kernel agrees."""
    g = torch.Generator().manual_seed(kind * 77 + cin + h)
    xb = torch.randn(b, h, w, cin, generator=g).to(dtype).cuda()
    if kind == 2:
        wt = (torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** 0.5).cuda()
        ref = F.conv_transpose2d(xb.float().permute(0, 3, 1, 2).double(), wt.to(dtype).double(), stride=2, padding=1)
    else:
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).cuda()
        ref = F.conv2d(xb.float().permute(0, 3, 1, 2).double(), wt.to(dtype).double(), padding=1)
    bias = torch.randn(cout, generator=g).cuda()
    ref = ref + bias.double()[None, :, None, None]
    kw = {}
    if film:
        sc, sh = (1 + 0.3 * torch.randn(b, cout, generator=g)).cuda(), torch.randn(b, cout, generator=g).cuda()
        kw.update(film_scale1p=sc, film_shift=sh)
        ref = ref * sc.double()[:, :, None, None] + sh.double()[:, :, None, None]
    r16 = (3.0 * torch.randn(b, ref.shape[2], ref.shape[3], cout, generator=g)).to(dtype).cuda()
    ref = ref + r16.double().permute(0, 3, 1, 2)
    wp = ops.pack_conv_weight(wt, kind, dtype)
    gn = 8 if cout % 32 == 0 else 0
    o = ops.conv_igemm(xb, wp, kind, cout, bias, resid=r16.clone(), want_f32=False, want_op=True, gn_groups=gn, **kw)
    d = ops.conv_direct(xb, wp, kind, cout, bias, resid=r16.clone(), want_f32=False, want_op=True, **kw)
    scale = max(float(ref.abs().max()), 1.0)
    half_ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    y = o["op"].double().permute(0, 3, 1, 2)
    assert o["op"].dtype == dtype and float((y - ref).abs().max()) < 1.2 * half_ulp * scale
    # the cross-check kernel rounds the same fp32 value (summation order aside: allow one 16-bit ulp on rare ties)
    assert float((o["op"].double() - d["op"].double()).abs().max()) <= 2.4 * half_ulp * scale
    assert float((o["op"] != d["op"]).double().mean()) < 1e-2
    if gn:
        rg = ref.reshape(b, gn, -1)
        mean, var = rg.mean(-1), rg.var(-1, unbiased=False)
        assert float((o["gn_stats"][..., 0].double() - mean).abs().max()) < 1e-4 * scale
        assert float((o["gn_stats"][..., 1].double() * (var + 1e-5).sqrt() - 1).abs().max()) < 1e-4
    # in place: the residual tensor itself receives the sum
    acc = r16.clone()
    o2 = ops.conv_igemm(xb, wp, kind, cout, bias, resid=acc, want_f32=False, want_op=True, inplace=True, **kw)
    assert o2["op"].data_ptr() == acc.data_ptr() and torch.equal(acc, o["op"])


@pytest.mark.parametrize("kind,b,h,w,cin,cout,resid", [
    (0, 2, 32, 32, 128, 128, True),      # cpg 16: two groups per 32-column chunk
    (0, 3, 16, 16, 32, 32, False),       # cpg 4
    (0, 2, 16, 16, 64, 64, True),        # cpg 8
    (0, 2, 16, 16, 256, 256, False),     # cpg 32: one group per chunk, two N... one N tile
    (0, 2, 8, 8, 512, 512, True),        # cpg 64: two chunks per group, two N tiles, half-filled M tile
    (1, 2, 32, 32, 128, 256, False),     # stride-2 producer
    (2, 2, 16, 16, 256, 128, True),      # transposed-conv producer (4 phases)
    (0, 1, 24, 24, 64, 64, False),       # ragged tiles: masked rows must not contribute
    (0, 3, 5, 256, 128, 128, True),      # row-slab mainloop on CTA pairs, odd tile count, per-warp residual sub-boxes
    (0, 1, 7, 40, 64, 64, False),        # 32-pixel tiles on a 40-pixel row: masked columns must not contribute
])
def test_conv_fused_groupnorm_statistics(ops, kind, b, h, w, cin, cout, resid):
    """The conv epilogue's fused (mean, rstd) equal a GroupNorm statistics pass over its own fp32 output."""
    g = torch.Generator().manual_seed(7 + cout + h)
    xb = torch.randn(b, h, w, cin, generator=g).to(torch.float16).cuda()
    wt = ((torch.randn(cin, cout, 4, 4, generator=g) if kind == 2 else torch.randn(cout, cin, 3, 3, generator=g)) * 0.05).cuda()
    bias = (torch.randn(cout, generator=g) + 0.5).cuda()
    oh, ow = (h // 2, w // 2) if kind == 1 else ((2 * h, 2 * w) if kind == 2 else (h, w))
    kw = {"resid": torch.randn(b, oh, ow, cout, generator=g).cuda()} if resid else {}
    o = ops.conv_igemm(xb, ops.pack_conv_weight(wt, kind), kind, cout, bias, gn_groups=8, **kw)
    y = o["f32"].double().reshape(b, oh * ow, 8, cout // 8)
    mean = y.mean(dim=(1, 3))
    rstd = 1.0 / torch.sqrt(y.var(dim=(1, 3), unbiased=False) + 1e-5)
    np.testing.assert_allclose(o["gn_stats"][..., 0].cpu().numpy(), mean.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o["gn_stats"][..., 1].cpu().numpy(), rstd.cpu().numpy(), rtol=1e-4)
    o2 = ops.conv_igemm(xb, ops.pack_conv_weight(wt, kind), kind, cout, bias, gn_groups=8, **kw)
    assert torch.equal(o["gn_stats"], o2["gn_stats"]) and torch.equal(o["f32"], o2["f32"])   # deterministic
    # finalize + apply == the standalone GroupNorm kernel on the same tensor
    gamma, beta = (1 + 0.1 * torch.randn(cout, generator=g)).cuda(), (0.1 * torch.randn(cout, generator=g)).cuda()
    a = ops.groupnorm_apply(o["f32"], gamma, beta, o["gn_stats"], 8, silu=True)
    ref = ops.groupnorm_silu(o["f32"], gamma, beta, 8, 1e-5, True)
    assert float((a.float() - ref.float()).abs().max()) <= 4e-3 * max(1.0, float(ref.float().abs().max()))


@pytest.mark.parametrize("b,h,w,cin,cout,silu,resid", [
    (1, 4, 256, 128, 128, True, False),    # CTA-pair slab mainloop, two channel blocks
    (3, 5, 192, 64, 64, True, True),       # ragged W (clipped second tile), odd tile count (masked pair half), residual
    (2, 6, 320, 128, 3, False, False),     # single-CTA slab, narrow N: the head (out_norm has no activation)
    (1, 3, 128, 192, 128, True, True),     # three channel blocks, W == one tile
    (8, 16, 128, 128, 128, True, True),    # many tiles per CTA: image changes inside a CTA's tile walk (table refresh)
])
def test_conv_fused_input_groupnorm(ops, b, h, w, cin, cout, silu, resid):
    """clpk_conv_epilogue.in_scale / in_shift: the conv applies act(x * scale[b,c] + shift[b,c]) to its own A operand in
    shared memory (GroupNorm apply + SiLU of the consumer side).  Reference: the same kernel WITHOUT the transform, fed
    with the transformed tensor computed by torch and rounded to fp16 — identical up to the 1-ulp disagreements between
    tanh.approx-based SiLU and torch's; and the fp64 convolution of that tensor.  Zero padding must stay zero."""
    g = torch.Generator().manual_seed(11 + cin + cout + w)
    x16 = (torch.randn(b, h, w, cin, generator=g) * 2 + 0.5).to(torch.float16).cuda()
    sc = (1 + 0.3 * torch.randn(b, cin, generator=g)).cuda()
    sh = (0.5 * torch.randn(b, cin, generator=g) + 0.7).cuda()      # non-zero shift: padding would show up as act(shift)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    assert ops.conv_in_affine_supported(0, h, w, cin, cout)
    t = x16.float() * sc[:, None, None, :] + sh[:, None, None, :]
    t16 = (F.silu(t) if silu else t).to(torch.float16)
    kw = {"resid": torch.randn(b, h, w, cout, generator=g).cuda()} if resid else {}
    nchw = cout % 16 != 0
    wp = ops.pack_conv_weight(wt, 0)
    args = dict(want_f32=not nchw, want_op=False, want_nchw=nchw, **kw)
    fused = ops.conv_igemm(x16, wp, 0, cout, bias, in_affine=(sc, sh, silu), **args)
    plain = ops.conv_igemm(t16, wp, 0, cout, bias, **args)
    y = fused["nchw"] if nchw else fused["f32"].permute(0, 3, 1, 2)
    yp = plain["nchw"] if nchw else plain["f32"].permute(0, 3, 1, 2)
    ref = F.conv2d(t16.float().permute(0, 3, 1, 2).double(), wt.to(torch.float16).double(), bias.double(), padding=1)
    if resid:
        ref = ref + kw["resid"].permute(0, 3, 1, 2).double()
    scale = max(1.0, float(ref.abs().max()))
    assert float(torch.linalg.norm(y.double() - ref) / torch.linalg.norm(ref)) < 1e-3
    assert float((y - yp).abs().max()) < 5e-3 * scale
    assert float((y.double() - ref).abs().max()) < 5e-3 * scale
    again = ops.conv_igemm(x16, wp, 0, cout, bias, in_affine=(sc, sh, silu), **args)
    assert torch.equal(again["nchw"] if nchw else again["f32"], fused["nchw"] if nchw else fused["f32"])   # deterministic
    assert not ops.conv_in_affine_supported(0, 32, 32, 128, 128) and not ops.conv_in_affine_supported(1, 256, 256, 128, 128)


def test_groupnorm_affine_from_conv_statistics(ops):
    """conv (fused statistics) -> clpk_groupnorm_affine -> next conv normalising its own operand  ==  conv -> GroupNorm+SiLU
    pass -> conv: the two-kernel ResBlock tail (blocks.py:43) without the stand-alone normalisation pass."""
    g = torch.Generator().manual_seed(21)
    b, h, w, c = 2, 6, 256, 128
    x16 = torch.randn(b, h, w, c, generator=g).to(torch.float16).cuda()
    w1 = (torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).cuda()
    w2 = (torch.randn(c, c, 3, 3, generator=g) / (c * 9) ** 0.5).cuda()
    b1, b2 = (torch.randn(c, generator=g) + 3.0).cuda(), torch.randn(c, generator=g).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(c, generator=g)).cuda(), (0.1 * torch.randn(c, generator=g)).cuda()
    o1 = ops.conv_igemm(x16, ops.pack_conv_weight(w1, 0), 0, c, b1, want_f32=True, want_op=True, gn_groups=8, return_partial=True)
    sc, sh = ops.groupnorm_affine(o1["gn_partial"], b, o1["gn_slots"], gamma, beta, 8)
    mean, rstd = o1["gn_stats"][..., 0], o1["gn_stats"][..., 1]
    np.testing.assert_allclose(sc.cpu().numpy(), (rstd.repeat_interleave(c // 8, dim=1) * gamma).cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(sh.cpu().numpy(), (beta - mean.repeat_interleave(c // 8, dim=1) * rstd.repeat_interleave(c // 8, dim=1) * gamma).cpu().numpy(),
                               rtol=1e-5, atol=1e-6)
    fused = ops.conv_igemm(o1["op"], ops.pack_conv_weight(w2, 0), 0, c, b2, in_affine=(sc, sh, True))["f32"]
    t16 = ops.groupnorm_apply(o1["op"].float(), gamma, beta, o1["gn_stats"], 8, silu=True)
    plain = ops.conv_igemm(t16, ops.pack_conv_weight(w2, 0), 0, c, b2)["f32"]
    assert float(torch.linalg.norm(fused - plain) / torch.linalg.norm(plain)) < 1e-3


@pytest.mark.parametrize("b,h,w,cin,cout,film,resid", [
    (2, 6, 256, 128, 128, True, False),     # conv1-like: FiLM, 16-bit output, statistics
    (3, 5, 192, 64, 64, False, True),       # odd H (masked second row of the last pair), ragged W, residual
    (1, 2, 128, 192, 128, False, True),     # three channel blocks, a single row pair
])
def test_conv_two_rows_per_item_variant(ops, monkeypatch, b, h, w, cin, cout, film, resid):
    """CLPK_IGEMM_ROWS2=2 (opt-in mainloop: two output rows per work item, separate slab / weight rings, four TMEM
    accumulators) against the default one-row slab mainloop: same products, only the fp32 accumulation order differs
    (channel block outer / kernel row inner instead of the reverse)."""
    g = torch.Generator().manual_seed(31 + h + cin)
    xb = torch.randn(b, h, w, cin, generator=g).to(torch.float16).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    kw = {}
    if film:
        kw.update(film_scale1p=(1 + 0.3 * torch.randn(b, cout, generator=g)).cuda(), film_shift=torch.randn(b, cout, generator=g).cuda())
    if resid:
        kw["resid"] = torch.randn(b, h, w, cout, generator=g).cuda()
    wp = ops.pack_conv_weight(wt, 0)
    ref = ops.conv_igemm(xb, wp, 0, cout, bias, want_f32=True, want_op=True, gn_groups=8, **kw)
    monkeypatch.setenv("CLPK_IGEMM_ROWS2", "2")
    two = ops.conv_igemm(xb, wp, 0, cout, bias, want_f32=True, want_op=True, gn_groups=8, **kw)
    monkeypatch.delenv("CLPK_IGEMM_ROWS2")
    scale = max(1.0, float(ref["f32"].abs().max()))
    assert float((two["f32"] - ref["f32"]).abs().max()) < 2e-5 * scale
    assert float((two["op"].float() - ref["op"].float()).abs().max()) <= 2.0 ** -10 * scale      # at most one fp16 ulp apart
    assert torch.allclose(two["gn_stats"], ref["gn_stats"], rtol=1e-5, atol=1e-6)
    assert not torch.equal(two["f32"], ref["f32"]) or cin == 64                                   # the switch changed the path


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("b,h,w,c", [
    (2, 6, 256, 128),      # two 128-pixel tiles per row, one row per CTA band (halo rows recomputed by the neighbours)
    (1, 5, 320, 64),       # ragged W: third tile half filled, one channel block
    (3, 9, 40, 128),       # W < 128: a single partial tile per row
    (2, 40, 128, 192),     # three channel blocks
    (8, 200, 256, 128),    # bands of several rows that cross image boundaries (ring slots of two images)
    (1, 1, 256, 128),      # single-row image: both vertical neighbours are padding
])
def test_head_conv_fused(ops, b, h, w, c, dtype):
    """clpk_head_conv (out_norm + 3x3 conv to 3 channels as pointwise tcgen05 GEMM + 9-point shift-add) against the fp64
    convolution of the normalised tensor rounded to the operand format, and against the implicit-GEMM kernel it replaces."""
    g = torch.Generator().manual_seed(5 + w + c + h)
    x16 = (torch.randn(b, h, w, c, generator=g) * 1.5 + 0.3).to(dtype).cuda()
    sc = (1 + 0.3 * torch.randn(b, c, generator=g)).cuda()
    sh = (0.5 * torch.randn(b, c, generator=g) + 0.4).cuda()
    wt = (torch.randn(3, c, 3, 3, generator=g) / (c * 9) ** 0.5).cuda()
    bias = torch.randn(3, generator=g).cuda()
    y = ops.head_conv(x16, sc, sh, wt, bias)
    assert y.shape == (b, 3, h, w) and torch.isfinite(y).all()
    t16 = (x16.float() * sc[:, None, None, :] + sh[:, None, None, :]).to(dtype)
    ref = F.conv2d(t16.float().permute(0, 3, 1, 2).double(), wt.to(dtype).double(), bias.double(), padding=1)
    scale = max(1.0, float(ref.abs().max()))
    ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    # the in-kernel FMA and torch's mul+add may round a normalised value to neighbouring 16-bit numbers
    assert float((y.double() - ref).abs().max()) < 8 * ulp * scale
    assert float(torch.linalg.norm(y.double() - ref) / torch.linalg.norm(ref)) < (1e-3 if dtype == torch.float16 else 6e-3)
    if w >= 128 and c % 64 == 0:
        old = ops.conv_igemm(t16, ops.pack_conv_weight(wt, 0, dtype), 0, 3, bias, want_f32=False, want_nchw=True)["nchw"]
        assert float((y - old).abs().max()) < 8 * ulp * scale
    assert torch.equal(y, ops.head_conv(x16, sc, sh, wt, bias))          # deterministic


@pytest.mark.parametrize("h,w,oh,ow", [(37, 53, 32, 32), (300, 200, 256, 256), (64, 64, 256, 256), (512, 768, 256, 256),
                                        (256, 256, 256, 256), (17, 9, 64, 48), (1000, 30, 31, 300), (256, 300, 256, 128)])
def test_bicubic_resize_bit_exact_vs_pillow(oracle, h, w, oh, ow):
    """The eval originals' BICUBIC resize on the device == Pillow's Image.resize byte for byte (down- and up-scaling, one
    pass skipped, extreme aspect ratios), and the float CHW conversion == numpy's of eval.py:67 bit for bit."""
    from PIL import Image
    from clip_neural_image_conpression_b200.eval.resample import resize_bicubic_u8, u8_hwc_to_float_chw
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    got = resize_bicubic_u8(cu(img), oh, ow)
    assert got.shape == (oh, ow, 3) and np.array_equal(got.cpu().numpy(), ref)
    assert np.array_equal(ref, oracle.bicubic_resize_u8(img, oh, ow))
    f = u8_hwc_to_float_chw(got).cpu().numpy()
    assert np.array_equal(f, oracle.original_to_float_chw(ref))


# ------------------------------------------------------------------------------------------------ blocks, post-process
def test_film_and_resblock_match_reference(ops, golden):
    from clip_neural_image_conpression_b200.models import FiLM, ResBlock
    g = golden("unet_forward")
    film = FiLM(16, 32)
    film.load_state_dict({k[len("film.sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("film.sd.")})
    y = film.cuda()(cu(g["film.x"]), cu(g["film.h"]))
    assert y.shape == (2, 16, 8, 8)                                        # the reference's own test (shape)
    np.testing.assert_allclose(y.cpu().numpy(), g["film.y"], rtol=1e-5, atol=1e-5)
    rb = ResBlock(32, 256)
    rb.load_state_dict({k[len("rb.sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("rb.sd.")})
    y = rb.cuda()(cu(g["rb.x"]), cu(g["rb.h"])).cpu()
    ref = torch.from_numpy(g["rb.y"])
    # fp16 conv operands, fp32 accumulation/stream: relative L2 far under the 1e-2 epsilon bar
    assert float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref)) < 1e-3
    rb.operand_dtype = torch.bfloat16
    yb = rb(cu(g["rb.x"]), cu(g["rb.h"])).cpu()
    assert float(torch.linalg.norm(yb - ref) / torch.linalg.norm(ref)) < 5e-3


def test_to_uint8_and_psnr(ops, oracle, golden):
    g = golden("metrics")
    u8 = ops.to_uint8_hwc(cu(g["a"])).cpu().numpy()
    assert np.array_equal(u8, g["u8"])                                      # bit exact (truncation)
    from clip_neural_image_conpression_b200.eval.metrics import psnr, psnr_batch
    got = psnr_batch(cu(g["a"]), cu(g["b"]))
    # the reference averages squares in float32 (numpy pairwise sum); ours is an exact integer sum -> ~1e-6 dB
    np.testing.assert_allclose(got, g["psnr"][:3], rtol=0, atol=1e-4)
    assert psnr(g["a"][0], g["a"][0]) == float("inf")
    sq = ops.psnr_sqerr_u8(cu(g["a"]), cu(g["b"])).cpu().numpy()
    ref = ((oracle.metric_uint8(g["a"]).astype(np.int64) - oracle.metric_uint8(g["b"]).astype(np.int64)) ** 2).reshape(3, -1).sum(1)
    assert np.array_equal(sq, ref)                                           # integer work: exact


def test_ssim_matches_oracle(ops, oracle, golden):
    """clpk_ssim_u8 vs the oracle's restatement of skimage.structural_similarity (float64 both sides)."""
    from clip_neural_image_conpression_b200.eval.metrics import ssim, ssim_batch
    g = golden("metrics")
    got = ssim_batch(cu(g["a"]), cu(g["b"]))
    ref = [oracle.ssim(g["a"][i], g["b"][i]) for i in range(len(got))]
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    rng = np.random.default_rng(11)
    for (c, h, w) in [(3, 7, 7), (3, 9, 41), (3, 37, 53), (2, 64, 64), (3, 256, 256)]:   # ragged tiles, minimum size
        a = rng.uniform(-1.2, 1.2, (2, c, h, w)).astype(np.float32)                        # values beyond [-1,1] clip
        b = (a + rng.normal(0, 0.3, a.shape)).astype(np.float32)
        got = ops.ssim_u8(cu(a), cu(b)).cpu().numpy()
        ref = []
        for i in range(2):
            planes = [oracle.ssim(np.repeat(a[i][k:k + 1], 3, 0), np.repeat(b[i][k:k + 1], 3, 0)) for k in range(c)]
            ref.append(np.mean(planes))
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9, err_msg=str((c, h, w)))
    assert ssim(g["a"][0], g["a"][0]) == 1.0
    assert abs(ssim(g["a"][0], g["b"][0]) - oracle.ssim(g["a"][0], g["b"][0])) < 1e-12
    with pytest.raises(ValueError):
        ops.ssim_u8(cu(g["a"][:, :, :5]), cu(g["b"][:, :, :5]))
    with pytest.raises(ValueError):
        ssim(g["a"][0][:1], g["b"][0][:1])                                  # (1,H,W): skimage raises, so do we


def test_ddpm_helpers_bit_exact(golden):
    """clpk_ddpm_combine through the NoiseScheduler mirror == the reference's CPU results, bit for bit (scheduler.py:46-68)."""
    from clip_neural_image_conpression_b200.diffusion import NoiseScheduler
    g = golden("ddpm")
    x0, noise, eps = (cu(g[k]) for k in ("x0", "noise", "eps"))
    t = torch.from_numpy(g["t"]).cuda()
    for sch in ("cosine", "linear"):
        s = NoiseScheduler(1000, sch, "cuda")
        xt = s.q_sample(x0, t, noise)
        assert np.array_equal(xt.cpu().numpy(), g[f"{sch}.xt"])
        assert np.array_equal(s.predict_x0_from_eps(xt, t, eps).cpu().numpy(), g[f"{sch}.x0_pred"])
        mean, var, x0c = s.p_mean_variance(lambda x, z, tt: eps, xt, None, t)
        assert np.array_equal(mean.cpu().numpy(), g[f"{sch}.mean"])
        assert np.array_equal(var.cpu().numpy(), g[f"{sch}.var"]) and var.shape == (4, 1, 1, 1)
        assert np.array_equal(x0c.cpu().numpy(), g[f"{sch}.x0_clamped"])
    # NaN propagates through the clamp like torch.clamp
    bad = x0.clone(); bad[0, 0, 0, 0] = float("nan")
    out = s.predict_x0_from_eps(bad, t, eps, clamp=True)
    assert torch.isnan(out[0, 0, 0, 0]) and float(out[1:].abs().max()) <= 1.0
