"""Instruction-level verification of the generated PTX Montgomery kernels: the generator's own
simulator executes the emitted instruction list against Python big integers (no GPU needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_ptx_simulates_correctly():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_field
    gen_field.selftest(iters=60, seed=3)


def test_header_is_up_to_date(tmp_path):
    out = tmp_path / "field_gen.cuh"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_field.py"), "--out", str(out)])
    cur = open(os.path.join(ROOT, "hekaton_system_b200", "csrc", "field_gen.cuh")).read()
    assert out.read_text() == cur, "field_gen.cuh is stale: run python tools/gen_field.py"
