"""Warp-stall breakdown of one kernel from an `ncu --set full --import-source on` report (run where ncu is installed; no GPU
needed): totals per stall reason, the share of samples on instructions only ONE warp per CTA executes (e.g. the single thread
that issues tcgen05.mma), and the hottest SASS lines.
usage: python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel-regex> [instructions-per-tile-count] > profiles/rN_stalls_<what>.txt"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
once = int(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) > 10]      # first launch only
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
n = lambda r, k: int(r[ix[k]] or 0)
tot = sum(n(r, "# Samples") for r in data)
print(f"report {rep}\nkernel {rows[hi[0] - 1][1][:120]}\nsampled warp-states {tot}, SASS instructions {len(data)}, "
      f"warp-instructions executed {sum(n(r, 'Instructions Executed') for r in data)}")
agg = {s: sum(n(r, s) for r in data) for s in stalls}
print("\nstall reason            samples   share")
for s, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v:
        print(f"  {s[6:]:20s}{v:9d}  {100 * v / tot:5.1f}%")
if once:
    one = sum(n(r, "# Samples") for r in data if n(r, "Instructions Executed") == once)
    warps = 8
    print(f"\ninstructions executed exactly {once} times (once per tile, by the one warp holding the MMA-issuing thread): "
          f"{one} samples = {100 * one / tot:.1f}% of all samples = {100 * warps * one / tot:.0f}% of that warp's time ({warps} warps per CTA)")
print("\nhottest SASS lines:  index  samples  executed  main stall  instruction")
for r in sorted(data, key=lambda r: -n(r, "# Samples"))[:40]:
    st = {s[6:]: n(r, s) for s in stalls if n(r, s)}
    print(f"  {data.index(r):5d} {n(r, '# Samples'):7d} {n(r, 'Instructions Executed'):9d}  {max(st, key=st.get) if st else '':14s} {r[ix['Source']].strip()[:90]}")
