"""bench.py contract pieces that run without a GPU: the CPU reference arm's JSON line and the algorithmic-byte count
 that `roofline.achieved` is built from."""
import json
import os
import subprocess
import sys

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--preroll", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("env steps/sec") and d["unit"] == "env steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "cologne8" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_algorithmic_bytes_follow_the_survey_formula():
    sys.path.insert(0, ROOT)
    import bench
    sc, m = util.marshal_map("cologne8")
    st = m.struct
    k = int(sum(len(p[0][1]) for p in m.info["programs_installed"].values()))
    vbar = 54.0
    expect = st.step_length * (vbar * 40.0 + st.n_lanes * 16.0 + st.n_signals * 16.0 + k) + st.n_sig_lanes * 20.0 \
        + st.n_signals * 56.0
    assert bench.algorithmic_bytes_per_step(m, vbar) == expect
    assert 100e3 < expect < 140e3          # ~121.6 KB per env step on cologne8
