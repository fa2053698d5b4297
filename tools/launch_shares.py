"""Kernel shares of a bench step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): total and mean
duration per kernel name.  usage: python tools/launch_shares.py launches.csv"""
import collections
import csv
import sys

tot, cnt = collections.Counter(), collections.Counter()
for row in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
    if row[0] == "ID" or len(row) < 15 or row[12] != "gpu__time_duration.sum":
        continue
    name = row[4].split("(")[0].replace("void ", "")
    name = name if len(name) < 70 else name[:67] + "..."
    ns = float(row[14].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row[13], 1)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v / 1e6:.3f} | {v / cnt[k] / 1e3:.1f} | {100 * v / total:.1f} % |")
