"""CPU: the C-ABI shared library loads, exports exactly the symbols include/b200dn.h declares, and fails
loudly (no fallback) when there is no sm_100 device."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest
import torch

from vub_image_denoising_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "b200dn.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    declared = _declared()
    assert declared == sorted(_lib.EXPORTS), "include/b200dn.h and _lib.EXPORTS disagree"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} missing from libb200dn.so"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (b200dn_[a-z0-9_]+)\b", out)))
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_abi_version_and_struct_size(built_lib):
    assert built_lib.b200dn_abi_version() == _lib.ABI_VERSION == 2
    # the ctypes mirror must have the C struct's size: 7 ints + pad, pointers ... computed by the C compiler
    src = '#include "b200dn.h"\n#include <stdio.h>\nint main(){printf("%zu", sizeof(b200dn_igemm_args));return 0;}'
    exe = ROOT / "vub_image_denoising_b200" / "build" / "sizeof_args"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", str(ROOT / "include"), "-o", str(exe)], input=src, text=True, check=True)
    size = int(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert ctypes.sizeof(_lib.IgemmArgs) == size


def test_igemm_args_field_offsets_match_the_header(built_lib):
    """Every field of the ctypes mirror sits where the C compiler puts it in struct b200dn_igemm_args."""
    c_names = ["mode", "prec", "B", "H", "W", "cin", "cout", "in", "in_ctot", "wpacked", "bias", "slope", "out_kind",
               "out", "out_ctot", "out_coff", "res", "res_ctot", "out_nchw", "res_nchw", "res_bmod", "block_n",
               "max_ctas", "m_tiles", "impl", "sat_flag"]
    body = "".join(f'printf("%zu ", offsetof(b200dn_igemm_args, {n}));' for n in c_names)
    src = f'#include "b200dn.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){{{body}return 0;}}'
    exe = ROOT / "vub_image_denoising_b200" / "build" / "offsets_args"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", str(ROOT / "include"), "-o", str(exe)], input=src, text=True, check=True)
    offs = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    py_names = [n for n, _ in _lib.IgemmArgs._fields_]
    assert [n.rstrip("_") for n in py_names] == c_names          # `in` is spelled `in_` in Python
    assert [getattr(_lib.IgemmArgs, n).offset for n in py_names] == offs


def test_dense_block_args_layout_matches_the_header(built_lib):
    """struct b200dn_dense_block_args: size and every field offset of the ctypes mirror vs the C compiler."""
    c_names = ["prec", "B", "H", "W", "channels", "in", "in_ctot", "out", "out_ctot", "out_coff", "wfused", "bias", "slope",
               "max_ctas", "sat_flag", "timeline"]
    body = "".join(f'printf("%zu ", offsetof(b200dn_dense_block_args, {n}));' for n in c_names)
    body += 'printf("%zu", sizeof(b200dn_dense_block_args));'
    src = f'#include "b200dn.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){{{body}return 0;}}'
    exe = ROOT / "vub_image_denoising_b200" / "build" / "offsets_dense"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", str(ROOT / "include"), "-o", str(exe)], input=src, text=True, check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    py_names = [n for n, _ in _lib.DenseBlockArgs._fields_]
    assert [n.rstrip("_") for n in py_names] == c_names
    assert [getattr(_lib.DenseBlockArgs, n).offset for n in py_names] == vals[:-1]
    assert ctypes.sizeof(_lib.DenseBlockArgs) == vals[-1]
    assert built_lib.b200dn_dense_block_weight_bytes(32) == 9 * (80 * 32 + 64 * 16 + 48 * 16 + 32 * 16) * 2
    assert built_lib.b200dn_dense_block_weight_bytes(64) < 0          # only the 32-channel block is fused
    a = _lib.DenseBlockArgs()
    h = ctypes.c_void_p(1)
    assert built_lib.b200dn_dense_block_prepare(ctypes.byref(a), ctypes.byref(h)) == -1 and h.value is None
    assert b"dense_block" in built_lib.b200dn_last_error()


def test_packed_weight_bytes(built_lib):
    f = built_lib.b200dn_packed_weight_bytes
    assert f(16, 16, 9, _lib.PREC_BF16) == 9 * 16 * 64 * 2          # cin padded to 64
    assert f(8, 24, 9, _lib.PREC_BF16) == 9 * 16 * 64 * 2           # cout padded to 16
    assert f(128, 320, 9, _lib.PREC_BF16X3) == 2 * 9 * 128 * 320 * 2
    assert f(0, 16, 9, 0) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_error_not_fallback(built_lib):
    a = _lib.IgemmArgs()
    rc = built_lib.b200dn_igemm(ctypes.byref(a), None)
    assert rc == -1 and b"igemm" in built_lib.b200dn_last_error()            # argument check first
    rc = built_lib.b200dn_sampler_step(8, 8, 8, 8, 0.0, 0.0, 0.0, 0.0, 8, 4, None)
    assert rc == -2, "without an sm_100 device compute calls must fail with B200DN_E_CUDA"
    with pytest.raises(RuntimeError):
        _lib.check(rc, "sampler_step")
    rc = built_lib.b200dn_igemm(None, None)
    assert rc == -1
    # prepared launches: argument errors before any device work, a failed prepare leaves no handle behind
    h = ctypes.c_void_p(123)
    assert built_lib.b200dn_igemm_prepare(ctypes.byref(a), ctypes.byref(h)) == -1 and h.value is None
    assert built_lib.b200dn_igemm_prepare(None, ctypes.byref(h)) == -1
    assert built_lib.b200dn_igemm_launch(None, None) == -1
    assert built_lib.b200dn_igemm_launch_list(None, 0, None) == -1
    assert built_lib.b200dn_igemm_rebind_nchw(None, None, None, 0) == -1
    built_lib.b200dn_igemm_release(None)          # releasing a null handle is a no-op
    # chained launches: same contract
    two = (_lib.IgemmArgs * 2)()
    assert built_lib.b200dn_conv_chain_workspace_bytes(None, 2) == -1
    two[0].B, two[0].H, two[0].W = 2, 32, 40
    assert built_lib.b200dn_conv_chain_workspace_bytes(two, 2) >= (2 * 2 * 5 + 2) * 4      # one counter per 8 x 16 tile
    h = ctypes.c_void_p(123)
    assert built_lib.b200dn_conv_chain_prepare(two, 1, 16, 0, ctypes.byref(h)) == -1 and h.value is None    # 2..4 layers
    assert built_lib.b200dn_conv_chain_prepare(two, 2, None, 0, ctypes.byref(h)) == -1                      # no workspace
    assert built_lib.b200dn_conv_chain_prepare(None, 2, 16, 0, ctypes.byref(h)) == -1


def test_stale_library_is_refused(built_lib, tmp_path, monkeypatch):
    """lib() compares the digest recorded at build time with the sources: an edited csrc/ must not run silently
    against the old binary (the .so is git-ignored and travels with the snapshot)."""
    from vub_image_denoising_b200 import _build
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_build, "_digest", lambda: "0" * 64)
    with pytest.raises(RuntimeError, match="stale"):
        _lib.lib()
    monkeypatch.setenv("B200DN_ALLOW_STALE", "1")
    assert _lib.lib() is not None


def test_modules_refuse_cpu_tensors():
    import vub_image_denoising_b200 as b2
    net = b2.RDUNet(base_filters=16).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 16, 16))
    dm = b2.DiffusionModel(b2.RDUNet_T(base_filters=16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dm.improved_sampling(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.metrics.calculate_psnr(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2.noise.add_gaussian_noise(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 25.0, seed=1)


def test_product_does_not_import_oracle():
    pkg = ROOT / "vub_image_denoising_b200"
    for py in pkg.rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{py} imports the oracle"
    for src in (pkg / "csrc").iterdir():
        assert "oracle/" not in src.read_text() or src.name == "philox_normal.h", f"{src} references oracle/"
    # developer tools are product-side too; studies that need the oracle live under tests/studies/
    for py in (ROOT / "tools").glob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", py.read_text(), flags=re.M), f"{py} imports the oracle"
    # bench.py: the oracle is imported in the CPU-baseline / reference-arm function only
    bench = (ROOT / "bench.py").read_text()
    imports = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, flags=re.M)]
    assert len(imports) == 1
    fn_start = bench.index("def cpu_reference_pass")
    fn_end = bench.index("\ndef ", fn_start + 1)
    assert fn_start < imports[0] < fn_end, "bench.py imports the oracle outside cpu_reference_pass"


def test_shim_aliases():
    import sys
    import vub_image_denoising_b200 as b2
    b2.shim.install()
    try:
        from UNet.RDUNet_model import RDUNet
        from diffusion_denoising.diffusion_RDUnet import RDUNet_T, DiffusionModel
        from diffusion_denoising.Unet.Unet_model import RDUNet_T as T2, init_weights
        assert RDUNet is b2.RDUNet and RDUNet_T is b2.RDUNet_T and T2 is b2.RDUNet_T
        assert DiffusionModel is b2.DiffusionModel and init_weights is b2.init_weights
    finally:
        b2.shim.uninstall()
    assert "UNet.RDUNet_model" not in sys.modules
