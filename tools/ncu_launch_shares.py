"""Aggregates an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`) into
per-kernel totals and SHARES of the summed device time -- the thing to compare with the `share_of_step` column of bench.py's
`kernels` table (ncu's per-launch times are cold-cache and serialised, so absolutes differ; the ranking and the shares must agree).
usage: python tools/ncu_launch_shares.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/rN_launch_shares.txt"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r is not hdr and r[ix["Metric Name"]] == "gpu__time_duration.sum"]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = data[skip:]
agg = {}
for r in data:
    name = re.sub(r"^(void\s+)?", "", r[ix["Kernel Name"]]).split("(")[0]
    name = name.replace("b200gs::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("tc5::", "")
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"{len(data)} launches, {tot / 1e3:.3f} ms of summed device time ({sys.argv[1]})")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {n:8d} {us:12.1f} {us / n:10.2f} {100 * us / tot:6.2f}%")
