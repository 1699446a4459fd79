#!/usr/bin/env python3
"""Offline view of what NVRTC compiles for a scene: writes /tmp/sdfk_jit_<scene>.cu (prelude + lowered body +
jit_kernels.cuh), compiles it with nvcc using the JIT flags and prints ptxas resource usage + an opcode histogram
of one kernel.   usage: python tools/dump_jit.py [scene] [kernel]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdfkit_b200 import scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "readme"
kernel = sys.argv[2] if len(sys.argv) > 2 else "sdfk_k_sample"
expr = {"readme": scenes.readme_scene, "csg50": scenes.csg50, "sphere": scenes.sphere, "perf": scenes.perf_scene}[name]()[0]
csrc = os.path.join(ROOT, "sdfkit_b200", "csrc")
src = open(os.path.join(csrc, "sdfk_prelude.h")).read()
from sdfkit_b200.exprs import lower  # noqa: E402
low = lower(expr, fast_div=(lambda c: True) if "--fastdiv" in sys.argv else None)
src += "\nSK_FN sk_float4 sdf_eval(sk_float3 p)\n{\n" + low.body + "\n}\n"
src += "SK_FN void sdf_eval2(sk_float3 p0, sk_float3 p1, sk_float4& r0, sk_float4& r1)\n{\n" + low.body2 + "\n}\n"
src += open(os.path.join(csrc, "jit_kernels.cuh")).read()
cu = "/tmp/sdfk_jit_%s.cu" % name
open(cu, "w").write(src)
cubin = cu.replace(".cu", ".cubin")
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-O3", "-std=c++17", "-fmad=false",
                    "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-lineinfo", "-Xptxas=-v", "-o", cubin, cu],
                   capture_output=True, text=True)
print("\n".join(l for l in r.stderr.splitlines() if "Used" in l or "Compiling" in l or "error" in l))
sass = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, cubin], capture_output=True, text=True).stdout
ops = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[m.group(1).split(".")[0]] += 1
print(kernel, "static SASS instructions:", sum(ops.values()))
print(ops.most_common(40))
open(cu.replace(".cu", ".sass"), "w").write(sass)
