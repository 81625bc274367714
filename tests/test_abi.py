"""The C-ABI library loads and exports every symbol include/llfe.h declares (CPU only: no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "llfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llfe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import low_level_feature_extraction_b200 as pkg

    lib = pkg.load_library()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"libllfe.so does not export {s}"
    # and the ctypes prototype table covers exactly the header
    assert sorted(pkg.PROTOTYPES) == syms


def test_version_and_error_string_without_gpu():
    import low_level_feature_extraction_b200 as pkg

    lib = pkg.load_library()
    assert lib.llfe_version() >= 100
    assert isinstance(lib.llfe_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch
    import low_level_feature_extraction_b200 as pkg

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.LlfeError) as e:
        pkg.Context(0)
    assert "no CPU fallback" in str(e.value) or "LLFE_E_NODEVICE" in str(e.value)
    with pytest.raises(RuntimeError):
        pkg.engine(0)


def test_product_code_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "low_level_feature_extraction_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
