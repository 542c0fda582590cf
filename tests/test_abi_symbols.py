"""The C-ABI library loads and exports every function include/hekaton_b200.h declares.
No compute call is made (there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hekaton_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hekaton_system_b200.backend import LIB_PATH
    if not os.path.exists(LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_device_is_an_error_not_a_fallback():
    """Without a GPU the context constructor must fail loudly."""
    import torch
    from hekaton_system_b200.backend import Context, HkError
    if torch.cuda.is_available():
        return
    try:
        Context("bls12_381")
    except HkError as e:
        assert e.code == -7
    else:
        raise AssertionError("Context() succeeded without a GPU")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hekaton_system_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
    # the build path as well: the field-kernel generator and the Rust binding generator take their constants from the
    # product's own tables, and the host mirror / ABI header never mention the oracle
    for rel in ("tools/gen_field.py", "tools/gen_rust_sys.py", "host/cp_groth16.hpp", "host/hk_node.cpp",
                "include/hekaton_b200.h"):
        txt = open(os.path.join(ROOT, rel)).read()
        assert "import oracle" not in txt and "from oracle" not in txt, rel


def test_generated_field_header_is_current_and_oracle_free():
    """Regenerating csrc/field_gen.cuh with the oracle package made unimportable gives the committed file: the
    kernels' constants come from hekaton_system_b200/curves.py alone."""
    import subprocess
    import sys
    import tempfile
    out = os.path.join(tempfile.mkdtemp(), "field_gen.cuh")
    code = ("import sys, runpy; sys.modules['oracle'] = None; sys.argv = ['gen_field.py', '--out', %r]; "
            "runpy.run_path(%r, run_name='__main__')" % (out, os.path.join(ROOT, "tools", "gen_field.py")))
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.DEVNULL)
    assert open(out).read() == open(os.path.join(ROOT, "hekaton_system_b200", "csrc", "field_gen.cuh")).read()


def test_default_build_uses_the_fused_dual_product():
    """hk_build_flags is a pure query (no CUDA call): the shipped library must have the measured-faster arithmetic
    variants on (fused dual product, Fq2 multiply through it) and the measured-slower Karatsuba multiply off."""
    from hekaton_system_b200 import backend as B
    if not os.path.exists(B.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    flags = B.load_library().hk_build_flags()
    assert flags & 1, "HK_BUILD_MUL2 missing"
    assert flags & 4, "HK_BUILD_FP2_DUAL missing"
    assert not flags & 2, "HK_BUILD_MUL_KARA must stay off by default"
