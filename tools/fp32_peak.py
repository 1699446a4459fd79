#!/usr/bin/env python3
"""Measures the non-FMA FP32 lane-op rate (FMUL + FADD, scalar and packed f32x2) of this GPU with tools/micro/f32x2
while sampling the SM clock, and writes profiles-ready JSON (read by bench.py like MEASURED_PEAKS.json).
usage: python tools/fp32_peak.py [out.json]   (tools/micro/f32x2 is built by sdfkit_b200/build.py's micro target)"""
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "tools", "micro", "f32x2")
src = exe + ".cu"
if not os.path.exists(exe):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-o", exe, src], check=True)
clk = []
stop = threading.Event()


def sample():
    while not stop.is_set():
        o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True).stdout.strip()
        if o:
            clk.append([float(x) for x in o.split(",")])
        stop.wait(0.05)


th = threading.Thread(target=sample, daemon=True)
th.start()
runs = []
for _ in range(3):
    out = subprocess.run([exe], capture_output=True, text=True).stdout
    runs.append(out)
    time.sleep(0.2)
stop.set()
th.join()
scalar = max(float(m) for o in runs for m in re.findall(r"scalar\s*:\s*[\d.]+ ms\s+([\d.e+]+) lane-op/s", o))
packed = max(float(m) for o in runs for m in re.findall(r"packed f32x2:\s*[\d.]+ ms\s+([\d.e+]+) lane-op/s", o))
sm = sorted(c[0] for c in clk)
res = {"fp32_nofma_lane_ops_per_s": scalar, "fp32x2_nofma_lane_ops_per_s": packed,
       "nominal_lane_ops_per_s": 148 * 128 * 1.965e9,
       "how": "tools/micro/f32x2.cu: 148*8 CTAs x 256 threads, 16 independent chains of FMUL+FADD (no FMA contraction), 20000 iterations, best of 3 runs, CUDA events",
       "clocks": {"sm_mhz_max_seen": max(sm) if sm else None, "sm_mhz_median": sm[len(sm) // 2] if sm else None,
                  "sm_max_mhz": max(c[1] for c in clk) if clk else None, "power_w_max": max(c[2] for c in clk) if clk else None, "samples": len(clk)},
       "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
