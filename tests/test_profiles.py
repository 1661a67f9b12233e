"""CPU-only: the committed evidence under profiles/ stays readable -- the bench lines carry every key of the bench
contract, the ncu exports parse with the tools that the README readings were made with."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(PROFILES, pattern)))
    assert files, pattern
    return files[-1]


def test_latest_bench_line_has_the_contract_keys():
    d = json.load(open(_latest("r0[0-9]_*_bench_1M.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(d["value"] - d["config"]["n_atoms"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] == 1 and c["value"] > 0
    assert d["gpu_launches"] >= 3 * d["steps"]  # K1, pair kernel, K3 per step at least
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = json.load(open(_latest("r0[0-9]_*_bench_reference.json")))
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_ncu_exports_parse():
    raw = _latest("r0[0-9]_*_force_tile_ncu_raw.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], capture_output=True, text=True)
    assert out.returncode == 0 and "k_force_tile" in out.stdout and "gpu__time_duration.sum" in out.stdout
    src = _latest("r0[0-9]_*_force_tile_ncu_source.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_profile.py"), src], capture_output=True, text=True)
    assert out.returncode == 0 and "LDS.128" in out.stdout and "by region between barriers" in out.stdout


def test_traffic_file_matches_the_ncu_export():
    import csv
    t = json.load(open(os.path.join(PROFILES, "force_kernel_traffic.json")))
    rows = list(csv.reader(open(os.path.join(ROOT, t["source"].split(" ")[0]), errors="ignore")))
    hdr = rows[0]
    rd = [float(r[hdr.index("dram__bytes_read.sum")]) for r in rows[2:]]
    wr = [float(r[hdr.index("dram__bytes_write.sum")]) for r in rows[2:]]
    assert any(abs(a * 1e6 - t["dram_bytes_read_per_launch"]) < 1e3 and abs(b * 1e6 - t["dram_bytes_write_per_launch"]) < 1e3
               for a, b in zip(rd, wr))


def test_cpu_baseline_table_tool_runs():
    """tests/cpu_baseline_table.py on the smallest size: the harness cell list and the reference's O(N^2)
    update_list(true) give the same number of pairs (asserted inside the tool) and a timing line comes out."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpu_baseline_table.py"), "--max-n", "1000",
                          "--rebuild-max-n", "1000", "--seconds", "0.2", "--pot", "lj,harm2d"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [json.loads(l) for l in out.stdout.strip().splitlines()]
    assert "host" in lines[0] and lines[0]["host"]["threads_used"] == 1
    assert [l["potential"] for l in lines[1:]] == ["lj", "harm2d"]
    for l in lines[1:]:
        assert l["atom_steps_per_s"] > 0 and l["update_list_s"] > 0 and l["pairs"] > 0
