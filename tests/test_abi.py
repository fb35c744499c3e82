"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/mudpt_b200.h
declares; the ctypes binding covers all of them; compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mudpt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mudpt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from mudpt_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} not exported"
        assert n in _lib.SIGNATURES, f"{n} missing from the ctypes binding"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.mudpt_abi_version() == 1


def test_no_link_dependency_on_torch_or_libcuda():
    import subprocess
    from mudpt_b200 import build
    out = subprocess.run(["ldd", build.lib_path()], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libcuda.so" not in out and "libcudart" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from mudpt_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(embed_dim=64, image_resolution=32, vision_layers=1, vision_width=128, vision_patch_size=16,
                      context_length=77, transformer_width=64, transformer_heads=1, transformer_layers=1, n_ctx=2,
                      prompt_depth=1, device=0)
    h = C.c_void_p()
    rc = lib.mudpt_create(C.byref(cfg), C.byref(h))
    assert rc < 0
    assert b"no CUDA device" in lib.mudpt_global_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA device"):
        from mudpt_b200.engine import Engine
        Engine({}, 2, 1, torch.device("cpu"))


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or the reference)."""
    pkg = os.path.join(ROOT, "mudpt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f
                assert "/root/reference" not in s, f
