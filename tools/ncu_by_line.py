"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line / function, using nvdisasm line info.
usage: python tools/ncu_by_line.py <src.csv> <library.so> <kernel-mangled-substring> [top_n]
(the library must be the binary that was profiled)"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# offset -> (file, line) for the kernel's section
off2line, cur, inside = {}, None, False
for ln in sass:
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        inside = (".text." in ln and kern in ln)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
# function ranges from the sources
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
funcs = {}
for fn in ["sim.cu", "sim_kernels.cuh"]:
    starts = []
    for i, ln in enumerate(open(os.path.join(ROOT, "resco_b200", "csrc", fn)), 1):
        m = re.match(r"^(?:RS_HEAVY|__device__|__global__|static|template|__host__)[^;]*?\b(\w+)\s*\([^;]*$", ln)
        if m and not ln.startswith("template <"):
            starts.append((i, m.group(1)))
    funcs[fn] = starts


def func_of(f, line):
    name = "?"
    for i, n in funcs.get(f, []):
        if i <= line:
            name = n
        else:
            break
    return name


rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
base = None
by_line = collections.defaultdict(lambda: collections.Counter())
by_func = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key = off2line.get(a - base) or ("?", 0)
    smp = float(r[ci["# Samples"]] or 0)
    ins = float(r[ci["Instructions Executed"]] or 0)
    thr = float(r[ci["Thread Instructions Executed"]] or 0)
    for d in (by_line[key], by_func[(key[0], func_of(*key))], tot):
        d["samples"] += smp; d["inst"] += ins; d["thread_inst"] += thr
        for s in stalls:
            d[s] += float(r[ci[s]] or 0)
print("total samples %.0f, warp instructions %.0f, avg active threads %.1f" % (tot["samples"], tot["inst"], tot["thread_inst"] / max(tot["inst"], 1)))
print("stall mix:", ", ".join("%s %.1f%%" % (s[6:], 100 * tot[s] / tot["samples"]) for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
print("\nby function (share of samples | share of warp instructions | avg active threads | top stalls)")
for k, d in sorted(by_func.items(), key=lambda kv: -kv[1]["samples"])[:25]:
    top = sorted(stalls, key=lambda s: -d[s])[:3]
    print("  %5.1f%% | %5.1f%% | %4.1f | %-16s %-18s %s" % (100 * d["samples"] / tot["samples"], 100 * d["inst"] / tot["inst"], d["thread_inst"] / max(d["inst"], 1), k[0], k[1],
                                                     " ".join("%s %.0f%%" % (s[6:], 100 * d[s] / max(d["samples"], 1)) for s in top)))
print("\nby line")
for k, d in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
    top = sorted(stalls, key=lambda s: -d[s])[:2]
    text = ""
    try:
        text = open(os.path.join(ROOT, "resco_b200", "csrc", k[0])).read().splitlines()[k[1] - 1].strip()[:90]
    except Exception:
        pass
    print("  %5.1f%% | %5.1f%% | %4.1f | %s:%d %s | %s" % (100 * d["samples"] / tot["samples"], 100 * d["inst"] / tot["inst"], d["thread_inst"] / max(d["inst"], 1), k[0], k[1],
                                                     " ".join("%s %.0f%%" % (s[6:], 100 * d[s] / max(d["samples"], 1)) for s in top), text))
