"""CPU fuzz of the run classification behind the one-launch border resolution of the MSM fold
(hekaton_system_b200/csrc/fold_runs.cuh: the same source the device kernels compile).  No GPU, no oracle: the check is
a brute-force grouping of the list's pieces (tests/host/fold_runs_fuzz.cpp)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fold_runs_matches_brute_force(tmp_path):
    exe = str(tmp_path / "fold_runs_fuzz")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "hekaton_system_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "fold_runs_fuzz.cpp"), "-o", exe])
    out = subprocess.run([exe, "40000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.match(r"ok: (\d+) lists resolved \((\d+) runs, (\d+) merges\), (\d+) lists rejected", out.stdout)
    assert m, out.stdout
    resolved, runs, merges, rejected = map(int, m.groups())
    assert resolved > 5000 and merges > 5000 and rejected > 5000, out.stdout      # both outcomes well covered
