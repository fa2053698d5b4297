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


def test_host_wave_agent_matches_numpy_rule_and_rejects_bad_tables():
    """rs_host_agent_wave is plain host code (no device, no RsSim): MAXWAVE over states.wave rows (skip = 0) against the
    numpy restatement of agents/maxwave.py:18-38, and the error path for a pair index outside the observation row."""
    import ctypes as C
    import numpy as np
    import util
    from resco_b200.sim import HostWaveAgent, load_library
    sc, m = util.marshal_map("cologne8")
    sig = m.info["signal_ids"]
    rng = np.random.default_rng(0)
    wave = rng.integers(0, 6, (257, len(sig), 12)).astype(np.float32)
    agent = HostWaveAgent(sc.meta["phase_pairs"], sc.meta["valid_acts"], sig, use_wave=True)
    ref = util.maxpressure_actions(sc, m, np.concatenate([np.zeros_like(wave[:, :, :1]), wave], 2))
    np.testing.assert_array_equal(agent.act(wave), ref)
    lib = load_library()
    out = np.zeros((1, len(sig)), np.int32)
    rc = lib.rs_host_agent_wave(wave.ctypes.data, 1, len(sig), 5, 0, agent.pairs.ctypes.data, agent.n_pairs,
                                agent.order.ctypes.data, out.ctypes.data)
    assert rc < 0 and b"outside the observation row" in lib.rs_last_error()
