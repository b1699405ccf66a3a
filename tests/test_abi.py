"""CPU: the C-ABI library loads and exports every symbol include/w2s.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from shap_transformer_asr_b200 import _lib
from shap_transformer_asr_b200.config import MODELS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        entry.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "w2s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2s_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_entry_points():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), f"libw2s.so does not export {name}"


def test_library_has_no_torch_or_cuda_driver_link_dependency():
    # plain C ABI: only libstdc++/libc style dependencies (cudart is linked statically)
    out = os.popen(f"ldd {_lib.LIB_PATH}").read()
    assert "torch" not in out and "libcuda.so" not in out


def test_config_struct_matches_header_layout():
    # 3 ints + 3*8 arrays + ... : keep ctypes mirror in sync with the C struct (all 4-byte fields)
    text = open(os.path.join(ROOT, "include", "w2s.h")).read()
    body = text[text.index("typedef struct {"): text.index("} w2s_config;")]
    n_arrays = len(re.findall(r"\[W2S_MAX_CONV_LAYERS\]", body))
    n_scalars = len(re.findall(r"^\s*(int32_t|float)\s+\w+;", body, flags=re.M))
    assert ctypes.sizeof(_lib.W2SConfig) == 4 * (n_scalars + 8 * n_arrays)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu(lib):
    cfg = _lib.W2SConfig()
    handle = ctypes.c_void_p()
    rc = lib.w2s_create(ctypes.byref(cfg), None, None, None, 0, 0, ctypes.byref(handle))
    assert rc != 0 and b"no CUDA device" in lib.w2s_last_error(None)
    from shap_transformer_asr_b200 import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine({}, MODELS["wav2vec2-tiny"])


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package (nor the B200 arm's helpers) may import it."""
    import glob
    pkg = os.path.join(ROOT, "shap_transformer_asr_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + glob.glob(os.path.join(pkg, "csrc", "*")):
        if os.path.isdir(path) or path.endswith((".o", ".so")):
            continue
        text = open(path, errors="ignore").read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "oracle/" not in text or path.endswith(".md"), path
    bench = open(os.path.join(ROOT, "bench.py")).read()
    body = bench[bench.index("def run_b200"):bench.index("def main")]
    uses = [m.start() for m in re.finditer(r"from oracle", body)]
    assert len(uses) == 1 and "CPU baseline" in body[uses[0] - 400:uses[0]]     # only inside the cpu_baseline leg


def test_hot_kernels_are_tcgen05_tma_code_issued_without_elect_loops(lib):
    """The shipped SASS is sm_100a tensor-core / TMA code (not a CUDA-core fallback), and no tcgen05 / TMA instruction
    sits in the ELECT + BRA.U.ANY loop ptxas emits when such code is guarded by `lane == 0` instead of elect.sync."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "HMMA"):
        assert mnemonic in sass, f"{mnemonic} missing from libw2s.so"
    assert sass.count("BRA.U.ANY") == 0
