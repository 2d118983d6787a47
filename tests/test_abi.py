"""CPU: the C-ABI library builds, loads and exports every symbol include/clc_b200.h declares;
the ctypes prototype table mirrors the header; ops refuse CPU tensors (no fallback)."""
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT


def _header_symbols(debug=False):
    """Symbols the header declares: the production ABI, or (debug=True) the CLC_DEBUG_ABI-only section."""
    src = open(os.path.join(ROOT, "include", "clc_b200.h")).read()
    m = re.search(r"#ifdef CLC_DEBUG_ABI(.*?)#endif /\* CLC_DEBUG_ABI \*/", src, re.S)
    part = m.group(1) if debug else src.replace(m.group(0), "")
    return sorted(set(re.findall(r"CLC_API\s+[\w\s\*]+?\b(clc_\w+)\s*\(", part)))


def test_build_and_exports():
    import __graft_entry__ as g
    g.build()
    from clc_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (clc_\w+)", out))
    declared = _header_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = [s for s in exported if s not in declared]
    assert not extra, f"exported but not declared: {extra}"
    assert not [s for s in exported if "debug" in s], "the production library must not export bring-up hooks"
    # the bring-up build exports the production ABI plus exactly the CLC_DEBUG_ABI section of the header
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.DEBUG_LIB_PATH], capture_output=True, text=True).stdout
    exported_dbg = set(re.findall(r" T (clc_\w+)", out))
    assert exported_dbg == set(declared) | set(_header_symbols(debug=True))
    assert sorted(_lib.DEBUG_PROTOTYPES) == _header_symbols(debug=True)


def test_prototypes_match_header_and_load():
    from clc_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _header_symbols()
    h = _lib.lib()  # resolves every symbol, sets argtypes
    assert h.clc_version() == 2
    assert h.clc_strerror(0) == b"ok"
    assert b"invalid" in h.clc_strerror(-1)


def test_sass_is_sm100a():
    from clc_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback():
    import clc_b200
    gc = clc_b200.GaussianConditional(None)
    y = torch.zeros(1, 4, 2, 2)
    with pytest.raises(TypeError, match="CUDA float32"):
        gc(y, torch.ones_like(y), y, training=False)
    with pytest.raises(TypeError, match="CUDA float32"):
        clc_b200.L2_or_pearson_corr(torch.zeros(2, 4, 2, 2), torch.zeros(1, 4, 4, 4), 2, 2)
    with pytest.raises(ValueError):
        gc.quantize(y, "bogus")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "clc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "compressai" not in re.sub(r"#.*|\"\"\"[\s\S]*?\"\"\"|//.*", "", src), f
