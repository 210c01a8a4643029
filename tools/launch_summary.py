"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv <cmd>`) by kernel: launches, total time, share.
    python tools/launch_summary.py gpurun_out/x.csv "description of the command" > profiles/x_summary.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    for line in f:
        if line.startswith('"ID"'):
            rows = [line]
        elif rows:
            rows.append(line)
r = list(csv.reader(rows))
hdr = r[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for x in r[1:]:
    if len(x) <= iv:
        continue
    try:
        v = float(x[iv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", x[ik])).replace("void ", "").replace("dqgp::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
scale = 1e6 if r[1][iu] in ("ns", "nsecond") else 1e3
print(f"ncu launch list of: {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print(f"(cold-cache, serialised launches: compare SHARES, not absolutes)  total {tot / scale:.1f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda t: -t[1][1])[:16]:
    print(f"{k:44s} {v[0]:6d} launches {v[1] / scale:9.2f} ms {100 * v[1] / tot:5.1f}%")
