"""The committed measurement artifacts stay readable by the tools that the summaries quote (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_launch_list_segments_into_one_bench_step():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_breakdown.py"),
                          os.path.join(PROF, "r02_launches_ncu.csv")], capture_output=True, text=True, check=True).stdout
    head = out.splitlines()[0]
    assert "118 launches" in head, head                       # one commitment batch + one proof batch
    rows = [ln.split("|") for ln in out.splitlines() if ln.startswith("| k_")]
    shares = {r[1].strip(): float(r[4].strip().rstrip(" %")) for r in rows}
    assert abs(sum(shares.values()) - 100.0) < 0.5
    assert max(shares, key=shares.get) == "k_msm_accum"       # the kernel the roofline is quoted for dominates


def test_bench_lines_carry_the_contract_keys():
    line = json.loads(open(os.path.join(PROF, "r02_bench_line.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["vs_baseline"] is None and line["gpu_launches"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-9
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    ref = json.loads(open(os.path.join(PROF, "r02_bench_line_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["config"] == line["config"] and ref["metric"] == line["metric"]
    for n in (2, 4, 8):
        multi = json.loads(open(os.path.join(PROF, f"r02_bench_{n}gpu.json")).read().strip().splitlines()[-1])
        assert multi["n_gpus"] == n and multi["scaling"] == "weak"
        assert 0.9 < multi["value"] / (n * line["value"]) < 1.1


def test_sweeps_are_all_checked():
    for curve in ("bls12_381", "bn254"):
        d = json.load(open(os.path.join(PROF, f"r02s_sweep_{curve}.json")))
        assert d["all_checked"] and len(d["rows"]) == 81
        for r in d["rows"]:
            assert r.get("exact") or r.get("checked"), r
