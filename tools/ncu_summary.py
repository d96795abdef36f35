"""Compact per-kernel summary of an ncu --set full report (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json profiles/ncu_dram_bytes.json] > profiles/rN_ncu_full_<what>.csv
--json writes {kernel name: mean dram__bytes_read.sum + dram__bytes_write.sum per launch} -- what bench.py reports as
`roofline.traffic` (it never carries a hand-typed constant)."""
import csv, json, re, subprocess, sys
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_elapsed.avg.per_second"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [w for w in WANT if w in idx]
w = csv.writer(sys.stdout)
w.writerow(cols); w.writerow([units[idx[c]] for c in cols])
for r in rows[2:]:
    w.writerow([r[idx[c]] for c in cols])
if "--json" in sys.argv:
    path = sys.argv[sys.argv.index("--json") + 1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = re.sub(r"^(void\s+)?(b200gs::)?(\(anonymous namespace\)::)?(tc5::)?", "", r[idx["Kernel Name"]].split("(")[0].split("<")[0]).strip()
        name = name.split("::")[-1]
        tot = 0.0
        for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[c]].replace(",", "")) * scale.get(units[idx[c]], 1.0)
        acc.setdefault(name, []).append(tot)
    json.dump({"source": f"ncu --set full --clock-control none, {rep.split('/')[-1]} (tools/gpu_ncu.sh), dram__bytes_read.sum + dram__bytes_write.sum, mean per launch",
               "kernels": {k: sum(v) / len(v) for k, v in acc.items()}, "launches": {k: len(v) for k, v in acc.items()}}, open(path, "w"), indent=1)
