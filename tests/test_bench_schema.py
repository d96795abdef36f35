"""CPU: the committed bench lines (profiles/*_bench_*.json, written by bench.py on a B200) carry every key the
measurement contract names. Guards the JSON schema; the numbers themselves are re-measured by the driver."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "clocks", "roofline", "gpu_launches"]


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        pytest.skip("no committed bench line")
    return json.load(open(files[-1])), files[-1]


def test_product_line_schema():
    d, path = _latest("r1*_bench_b200_n1.json")
    for k in BASE + ["cpu_baseline", "kernels"]:
        assert k in d, (k, path)
    assert d["metric"] == "train_iters_per_s" and d["unit"] == "view-iters/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not set(d["clocks"]["reasons"]) & {
        "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(sum(k["share_of_step"] for k in d["kernels"]) - 1.0) < 0.25          # entry points cover most of the step


def test_reference_line_schema():
    d, path = _latest("r1*_bench_reference.json")
    for k in BASE:
        assert k in d, (k, path)
    assert d["impl"] == "reference" and d["metric"] == "train_iters_per_s"


def test_scaling_lines_are_weak_scaling_of_the_same_workload():
    lines = [json.load(open(f)) for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r1i_bench_b200_n*.json")) +
                                                glob.glob(os.path.join(ROOT, "profiles", "r1j_bench_b200_n1.json")))]
    if len(lines) < 2:
        pytest.skip("no scaling lines")
    assert all(l["scaling"] == "weak" and l["config"]["views_per_gpu"] == lines[0]["config"]["views_per_gpu"] for l in lines)
    by_n = {l["n_gpus"]: l["value"] for l in lines}
    assert all(by_n[n] > 0.85 * n * by_n[1] for n in by_n)


def test_round2_lines_schema_and_strong_scaling_of_the_named_batch():
    """The final round-2 lines: BASELINE.json config 3 as written (global batch 8 sharded over N GPUs, strong scaling) with the
    weak block next to it, the reference arm of the same visit, and the extra blocks DESIGN.md §6 quotes."""
    d, path = _latest("r4q_bench_b200_n1.json")
    r, _ = _latest("r4q_bench_reference.json")
    for k in BASE + ["cpu_baseline", "kernels", "timeline", "render", "raster_only", "stress_c5", "launcher_path"]:
        assert k in d, (k, path)
    for k in BASE:
        assert k in r, k
    assert d["scaling"] == "strong" and d["config"]["global_batch"] == 8 and r["impl"] == "reference"
    assert d["config"]["kernel_options"]["sort_ballot_rank"] == 1 and d["config"]["kernel_options"]["mlp_bwd_v2"] == 87
    assert d["metric"] == r["metric"] and d["unit"] == r["unit"] and d["config"]["workload"] == r["config"]["workload"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 8 * 3 * 720 * 1280
    assert d["roofline"]["traffic"] and "ncu --set full" in d["roofline"]["traffic_source"]
    assert d["value"] > 10 * r["value"] and d["e2e"]["value"] > 10 * r["e2e"]["value"]
    assert d["cpu_baseline"]["c1"]["config"].startswith("C1")
    for tag in ("1920x1080", "1280x720"):
        assert d["render"][tag]["e2e_fps"] > 5 * r["render"][tag]["e2e_fps"]
    lines = {}
    for pat in ("r4o_bench_b200_n2.json", "r4n_bench_b200_n4_early_sh_tail.json", "r4r_bench_b200_n8.json"):
        l, _ = _latest(pat)
        lines[l["n_gpus"]] = l
    for n, l in lines.items():
        assert l["scaling"] == "strong" and l["config"]["global_batch"] == 8 and l["config"]["views_per_gpu"] == 8 // n
        assert l["weak"]["scaling"] == "weak" and l["weak"]["views_per_gpu"] == 8 and l["weak"]["global_batch"] == 8 * n
        assert l["value"] > d["value"] and l["weak"]["value"] > 0.9 * n * d["value"]
        assert not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
