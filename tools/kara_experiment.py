#!/usr/bin/env python3
"""Offline experiment: compile one generated multiply variant into a micro kernel and print its SASS pipe mix."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import gen_field as g

def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else ""
    d = os.path.join(ROOT, "hekaton_system_b200", "csrc", "build", "scratch")
    os.makedirs(d, exist_ok=True)
    name, p = g.field_list()[0]
    fg = g.FieldGen(name, p)
    if variant == "old":
        P, outs, ins = fg.gen_mul()
    elif variant == "sqr":
        P, outs, ins = fg.gen_sqr()
        ins = ins + [f"b{k}" for k in range(fg.N2)]
    else:
        P, outs, ins = fg.gen_mul_kara(variant)
    src = ("#include <cstdint>\nstruct alignas(16) F { uint64_t v[6]; };\n" +
           fg.cpp_fn("x", P, outs, ins).replace(f"{name}_t", "F").replace(f"{name}_x", "fmul") +
           "__global__ void k(F* o, int it) {\n  F a = o[threadIdx.x], b = o[threadIdx.x + 256];\n"
           "#pragma unroll 1\n  for (int i = 0; i < it; i++) { F r; fmul(r, a, b); b = a; a = r; }\n  o[threadIdx.x] = a;\n}\n")
    cu = os.path.join(d, "exp.cu")
    open(cu, "w").write(src)
    cubin = os.path.join(d, "exp.cubin")
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v",
                        "-cubin", "-o", cubin, cu] + sys.argv[2:], capture_output=True, text=True)
    print([l for l in r.stderr.splitlines() if "registers" in l or "spill" in l])
    subprocess.run([os.path.join(ROOT, "tools", "sass_mix.sh"), cubin, "_Z1kP1Fi", "1"])

main()
