"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ctvio.h declares, and refuses to run without a CUDA device (no silent CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from helpers import pkg, syn


def header_symbols():
    hdr = open(os.path.join(pkg.REPO_ROOT, "include", "ctvio.h")).read()
    return sorted(set(re.findall(r"\b(ctvio_[a-z_0-9]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    names = header_symbols()
    assert sorted("ctvio_" + n for n in pkg.ABI_SYMBOLS) == names


def test_library_exports_every_declared_symbol():
    lib = pkg.load()  # raises if libctvio_b200.so is missing or a symbol is absent
    raw = C.CDLL(lib.path)
    for name in header_symbols():
        assert hasattr(raw, name), name
    assert raw.ctvio_abi_version() == pkg.binding.ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(pkg.Config) == 8 * 2 + 8 * 4 + 8 * 3 + 8 + 8 * 3 + 8 * 6 + 8 + 8 + 8 + 4 + 4
    assert C.sizeof(pkg.Options) == 4 * 8 + 8 * 2
    assert C.sizeof(pkg.Summary) == 4 * 8 + 8 * 4 + 8


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the no-device path cannot be exercised")
    except ImportError:
        pass
    lib = pkg.load()
    w = syn.config_c1()
    with pytest.raises(pkg.CtvioError) as ei:
        pkg.Estimator(lib, pkg.make_config(**w.config_kwargs()))
    assert "-2" in str(ei.value) or "no CUDA device" in str(ei.value)


def test_product_does_not_link_or_import_the_oracle():
    # the product path must never route through oracle/ (only tests, smoke() and bench.py's cpu_baseline may)
    for root, _, files in os.walk(pkg.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".sh")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), os.path.join(root, f)
