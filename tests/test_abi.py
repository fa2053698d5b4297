"""The C-ABI library loads without a GPU and exports every symbol include/resco_b200.h declares;
the product path has no CPU fallback (rs_create must fail loudly without a device)."""
import ctypes
import os
import re

import pytest

import util
from resco_b200 import abi
from resco_b200.sim import LIB_PATH, EXPORTS, build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "resco_b200.h")).read()
    return sorted(set(re.findall(r"\b(rs_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    build_library()
    lib = ctypes.CDLL(LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/resco_b200.h but not exported"
    assert sorted(EXPORTS) == names
    assert lib.rs_abi_version() == abi.RS_ABI_VERSION


def test_struct_layout_matches_header():
    """Field order of the ctypes mirror == field order of the C struct (same names, same sequence)."""
    src = open(os.path.join(ROOT, "include", "resco_b200.h")).read()
    body = src[src.index("typedef struct RsScenario {"):src.index("} RsScenario;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = body[body.index("{") + 1:]
    fields = []
    for stmt in body.split(";"):
        decl = stmt.strip()
        if not decl:
            continue
        for ty in ("const", "int32_t", "float", "uint8_t"):
            decl = re.sub(r"\b" + ty + r"\b", " ", decl)
        fields += [x.strip(" *\n") for x in decl.split(",") if x.strip(" *\n")]
    assert fields == [n for n, _ in abi.RsScenario._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from resco_b200.sim import VecSim, RsError
    sc, m = util.marshal_map("cologne1")
    with pytest.raises(RsError) as e:
        VecSim(m, 1)
    assert "no CUDA device" in str(e.value) or "error -" in str(e.value)


def test_product_never_imports_oracle():
    """Nothing under resco_b200/ may import, load or link the CPU oracle."""
    pkg = os.path.join(ROOT, "resco_b200")
    bad = re.compile(r"(import\s+pyoracle|from\s+pyoracle|liboracle|microsim\.c|orc_[a-z_]+\s*\()")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), f"{f} references the oracle"
