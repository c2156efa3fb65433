"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol include/lokib200.h declares, and refuses to run
without a device (there is no CPU fallback)."""
import os
import re

import pytest

import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import loki_mc_b200 as lk
    from loki_mc_b200._capi import HOST_SYMBOLS, SYMBOLS
    if not os.path.exists(lk.lib_path()):
        lk.build()
    L = lk.lib()
    for header, symbols in (("lokib200.h", SYMBOLS), ("lokib200_host.h", HOST_SYMBOLS)):
        hdr = open(os.path.join(ROOT, "include", header)).read()
        declared = sorted(set(re.findall(r"\b(lokib200_[a-z0-9_]+)\s*\(", hdr)))
        assert declared, "no declarations found in " + header
        for s in declared:
            assert hasattr(L, s), "missing symbol " + s
        assert sorted(symbols) == declared, header
    assert L.lokib200_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    import loki_mc_b200 as lk
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lk.LokiB200Error, match="no usable CUDA device"):
        lk.Engine(gio.load("reid_dc"), 128)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "loki_mc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lokioracle" not in src and "oracle/" not in src, f


def test_no_reduction_result_on_a_live_uniform_register():
    """ptxas 12.9 once placed the uniform destination of a REDUX on top of the live shared-memory base in two instantiations of the
    advance kernel (wild LDS address, found with compute-sanitizer).  The SASS of the built library is scanned for that pattern."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_ur_clobber
    import loki_mc_b200 as lk
    assert check_ur_clobber.scan(lk.lib_path()) == []
    assert check_ur_clobber.scan_derived_bases(lk.lib_path()) == []
