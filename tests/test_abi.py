"""The C-ABI shared library: builds, loads, exports every symbol include/clpk.h declares, and refuses to compute
without a GPU (no CPU fallback exists).  CPU only — no compute call is made."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "clpk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(clpk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_typed(lib):
    from clip_neural_image_conpression_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clpk.h but not exported by libclpk.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with include/clpk.h"
    assert lib.clpk_version() >= 100
    assert _lib.lib_path().parent == ROOT / "clip_neural_image_conpression_b200" / "csrc"   # built in-tree


def test_struct_layouts_match_header(lib):
    from clip_neural_image_conpression_b200._lib import ConvEpilogue, UnetConfig
    assert C.sizeof(ConvEpilogue) == 120 and ConvEpilogue.cout_valid.offset == 64 and ConvEpilogue.gn_cpg.offset == 80
    assert ConvEpilogue.in_scale.offset == 88 and ConvEpilogue.in_silu.offset == 104 and ConvEpilogue.resid_op.offset == 112
    assert C.sizeof(UnetConfig) == (3 + 8 + 4) * 4


def test_library_is_sm100a_tcgen05(lib):
    """The conv kernel in the shipped .so really is the Blackwell tensor-core path (SASS mnemonics)."""
    import shutil
    import subprocess
    from clip_neural_image_conpression_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):   # tcgen05.mma, TMA tensor load, tcgen05.ld
        assert mnemonic in sass, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_anywhere():
    from clip_neural_image_conpression_b200 import ops
    from clip_neural_image_conpression_b200._lib import ClpkError
    from clip_neural_image_conpression_b200.codecs import PerChannelAffineQuantizer
    from clip_neural_image_conpression_b200.diffusion import DDIMSampler, NoiseScheduler
    from clip_neural_image_conpression_b200.models import CLIPCondUNet, FiLM, timestep_embedding
    net = CLIPCondUNet(z_dim=16, base=32, ch_mult=(1,))
    with pytest.raises(ClpkError):
        net(torch.randn(1, 3, 8, 8), torch.randn(1, 16), torch.zeros(1, dtype=torch.long))
    with pytest.raises(ClpkError):
        FiLM(16, 32)(torch.randn(2, 16, 8, 8), torch.randn(2, 32))
    with pytest.raises(ClpkError):
        timestep_embedding(torch.zeros(2, dtype=torch.long), 8)
    with pytest.raises(ClpkError):
        ops.dequant_l2norm(torch.zeros((1, 4), dtype=torch.uint8), torch.ones(4), torch.zeros(4))
    with pytest.raises(ClpkError):
        PerChannelAffineQuantizer().fit(torch.randn(4, 4))
    with pytest.raises(ClpkError):
        DDIMSampler(NoiseScheduler(device="cpu")).sample(net, torch.randn(1, 16), (1, 3, 8, 8), steps=2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_plan_create_reports_cuda_error_without_gpu(lib):
    from clip_neural_image_conpression_b200._lib import UnetConfig
    cfg = UnetConfig()
    cfg.z_dim, cfg.base, cfg.n_levels, cfg.time_dim, cfg.img_ch, cfg.groups = 16, 32, 1, 256, 3, 8
    cfg.ch_mult[0] = 1
    handle = C.c_void_p()
    names = (C.c_char_p * 1)(b"x")
    ptrs = (C.c_void_p * 1)(None)
    numels = (C.c_int64 * 1)(0)
    rc = lib.clpk_plan_create(C.byref(cfg), 1, 8, 8, 0, names, ptrs, numels, C.byref(handle))
    assert rc == 2 and b"failed" in lib.clpk_last_error()      # CLPK_ERR_CUDA, with a message
