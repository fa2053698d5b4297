"""profiles/kernel_counters.json from the committed ncu raw exports: per bench config, DRAM bytes and executed warp
instructions of ONE launch of the fused env-step kernel, divided by the instances that launch stepped.

usage: python tools/ncu_counters.py c2=gpurun_out/r02_c2_raw.csv:4096 c3=...:8192 ...
(raw.csv = `ncu -i X.ncu-rep --page raw --csv`; the number after the colon is the instances per launch)"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "kernel_counters.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=")
    path, n_env = rest.rsplit(":", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    ci = {n: i for i, n in enumerate(hdr)}
    SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}

    def num(name):      # in bytes / ns / plain counts whatever unit ncu chose to print
        return float(vals[ci[name]].replace(",", "")) * SCALE.get(units[ci[name]], 1.0)
    n = int(n_env)
    out[key] = dict(n_env_per_launch=n, kernel=vals[ci["Kernel Name"]],
                    dram_bytes_per_env_step=(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / n,
                    warp_insts_per_env_step=num("smsp__inst_executed.sum") / n,
                    threads_per_warp_inst=num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    issue_active_pct=num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    duration_ns_under_ncu=num("gpu__time_duration.sum"),
                    source=os.path.basename(path) + " (ncu --set full --clock-control none, one launch)")
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
