"""The Rust side of the drop-in boundary (rust/, patches/): no Rust toolchain exists in this image, so what CAN be
checked is checked -- the `-sys` crate is generated from the header and current, it binds exactly the symbols the
shared library exports, the safe wrapper only calls functions that exist (and does bind the fused path the
benchmark measures), and the reference-side patches apply to the reference tree when it is present."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS_RS = os.path.join(ROOT, "rust", "hekaton-b200-sys", "src", "lib.rs")
WRAP = os.path.join(ROOT, "rust", "hekaton-b200", "src")
HEADER = os.path.join(ROOT, "include", "hekaton_b200.h")
LIB = os.path.join(ROOT, "hekaton_system_b200", "lib", "libhekaton_b200.so")


def _sys_fns():
    return set(re.findall(r"pub fn (hk_\w+)\(", open(SYS_RS).read()))


def test_sys_crate_is_generated_from_the_header_and_current():
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"])
    assert rc == 0, "run tools/gen_rust_sys.py"


def test_sys_crate_binds_every_exported_symbol():
    import __graft_entry__ as entry
    entry.build()
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    header = set(re.findall(r"\b(hk_\w+)\s*\(", text))
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split()[-1].startswith("hk_")}
    fns = _sys_fns()
    assert fns == header, (fns ^ header)
    assert fns == exported, (fns ^ exported)
    # the struct that crosses the boundary by pointer has the C field order
    rs = open(SYS_RS).read()
    fields = re.findall(r"pub (\w+):", rs[rs.index("pub struct hk_pk_desc"):rs.index('#[link(name')])
    assert fields == ["a_g", "b_g", "h_g", "b_h", "ck", "num_stages", "deltas_g", "alpha_g", "beta_g", "beta_h",
                      "last_delta_h", "r1cs", "stage_start", "witness_density_permille"]


def test_wrapper_crate_calls_only_bound_functions_and_covers_the_fused_path():
    fns = _sys_fns()
    used = set()
    for f in os.listdir(WRAP):
        used |= set(re.findall(r"sys::(hk_\w+)\s*\(", open(os.path.join(WRAP, f)).read()))
    assert used <= fns, used - fns
    # what the verdict asked for: the fused path, context management, checked MSM, scalar_pairing, witness map
    for need in ("hk_pk_create", "hk_pk_free", "hk_commit", "hk_prove_batch", "hk_ctx_create", "hk_ctx_fork", "hk_ctx_sync",
                 "hk_ctx_destroy", "hk_bases_upload_g1", "hk_bases_upload_g2", "hk_bases_free", "hk_bases_precompute",
                 "hk_msm_g1", "hk_msm_g2", "hk_msm_g1_checked", "hk_scalar_mul_g1", "hk_scalar_mul_g2", "hk_r1cs_upload",
                 "hk_qap_witness_map", "hk_last_error_string"):
        assert need in used, need
    lib_rs = open(os.path.join(WRAP, "lib.rs")).read()
    for mod in ("context", "error", "fused", "msm", "pack", "reduction"):
        assert f"pub mod {mod};" in lib_rs and os.path.exists(os.path.join(WRAP, mod + ".rs"))
    red = open(os.path.join(WRAP, "reduction.rs")).read()
    assert "impl R1CSToQAP for GpuReduction" in red and "fn witness_map_from_matrices" in red


@pytest.mark.skipif(not os.path.isdir("/root/reference/cp-groth16"), reason="reference tree not present on this box")
def test_patches_apply_to_the_reference(tmp_path):
    if shutil.which("patch") is None:
        pytest.skip("no patch(1)")
    for sub in ("cp-groth16/src", "distributed-prover/src"):
        shutil.copytree(os.path.join("/root/reference", sub), tmp_path / sub)
    shutil.copy("/root/reference/distributed-prover/Cargo.toml", tmp_path / "distributed-prover" / "Cargo.toml")
    pdir = os.path.join(ROOT, "patches")
    for p in sorted(f for f in os.listdir(pdir) if f.endswith(".patch")):
        subprocess.check_call(["patch", "-p1", "-s", "-i", os.path.join(pdir, p)], cwd=tmp_path)
    assert "CPGroth16::<E, QAP>::prove_last_stage_with_zk" in (tmp_path / "cp-groth16/src/committer.rs").read_text()
    assert "hekaton_b200::GpuReduction as QAP" in (tmp_path / "distributed-prover/src/util.rs").read_text()
