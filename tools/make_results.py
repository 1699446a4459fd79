#!/usr/bin/env python3
"""Regenerates RESULTS.md from the measured JSON lines under profiles/.  usage: python tools/make_results.py [tag]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01s4"
late = sys.argv[2] if len(sys.argv) > 2 else "r01s5"        # a later 1-GPU run (kernels only got faster; its box had slower PCIe)
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    return json.load(open(path)) if os.path.exists(path) else None


d = load(tag + "_bench_1gpu.json")
last = load("r01s6_bench_1gpu.json")                          # the very last bench.py run of the session
multi = {n: (load("%s_bench_%dgpu.json" % (tag, n)) or load("r01s4_bench_%dgpu.json" % n)) for n in (2, 4, 8)}
cfg = [json.loads(l) for l in open(os.path.join(P, late + "_configs.txt")) if l.startswith("{")]
d5 = load(late + "_bench_1gpu.json")
st = d5["stages_ms"]
FP32 = 3.49e13


def bench_row(n, x):
    if x is None:
        return ""
    g = x["config"]["grid"][0]
    e = x["e2e"]
    return "| %d | %d³ | %.2f | %.3g | %.3g | %.2f ms = %.3g voxels/s | %.2f ms = %.3g voxels/s (%.0f MB D2H per step) |\n" % (
        n, g, x["ms_per_step"], x["value"], x["tris_per_s"], x["fused_to_mesh"]["ms_per_step"], x["fused_to_mesh"]["value"],
        e["ms_per_step"], e["value"], e["d2h_bytes_per_step"] / 1e6)


txt = """# RESULTS — round 1 (measured on boxes of the pool, NVIDIA B200, SM clock 1965 MHz, no throttle reasons)

All GPU numbers: CUDA events on the library's stream, after ≥ 3 warm-ups; parity mode (IEEE f32/f64, no FMA contraction),
every output bit-exact against the CPU oracle and the committed golden fixtures in the GPU test-suite (113 tests, incl.
512³ / 1024³ property tests and exhaustive 2^32 checks of the packed sqrt / constant division). CPU numbers: the C++
restatement of the reference's CPU path (`oracle/`, g++ -O2 -ffp-contract=off) on the box's 16 host cores, on a bounded
sample of the same workload — the .NET reference itself cannot run here. Raw lines and ncu summaries: `profiles/%s_*`
(`profiles/r01_*`, `r01s2_*` are earlier snapshots of the round). Roofline denominators: HBM 6553 GB/s (measured copy,
`MEASURED_PEAKS.json`); FP32 without FMA 3.49e13 lane-op/s (measured FMUL+FADD chains). Box-to-box spread: ±2 %% on kernels; the e2e
call follows the box's PCIe (observed over the session: 6.2 – 6.5 ms on most boxes, 7.5 ms on one). The table below is the `%s` snapshot
(all four GPU counts within the same hour); stage times and the configuration table are from the last run of the session (`%s`:
step %.2f ms, fused step %.2f ms, e2e %.2f ms on a box with slower PCIe). The last `bench.py` run of the session (`r01s6`, all changes
in): step %.2f ms = %.3g voxels/s, fused step %.2f ms, e2e %.2f ms = %.3g voxels/s, K1 at %.2f of the HBM peak. (regenerate with `python tools/make_results.py %s`)

## bench.py (README RepeatXY scene → Voxels (clip) → MarchingCubes; one step = sample + mesh)

| GPUs | grid | ms/step | voxels/s (whole job) | tris/s | fused `Sdf.ToMesh` step (device) | e2e `Sdf.ToMesh` (mesh in host memory) |
|---|---|---|---|---|---|---|
""" % (tag, tag, late, d5["ms_per_step"], d5["fused_to_mesh"]["ms_per_step"], d5["e2e"]["ms_per_step"],
       last["ms_per_step"], last["value"], last["fused_to_mesh"]["ms_per_step"], last["e2e"]["ms_per_step"], last["e2e"]["value"], last["roofline"]["frac"], tag)
txt += bench_row(1, d)
for n in (2, 4, 8):
    txt += bench_row(n, multi[n])
cb = d["cpu_baseline"]
txt += """
Weak scaling of the step: 1 → 8 GPUs = %.2f× (8-GPU line: %s). The multi-GPU e2e number is bounded by the host side of the box (a
32-vCPU, single-NUMA-node VM): 8 ranks streaming their shares concurrently reach ≈ 91 GB/s in total (937 MB in 10.3 ms), against
55 GB/s for one GPU alone; up to 4 GPUs every rank still gets its full link (e2e 5.9–6.4 ms).

CPU restatement on a 256³ sample (16 cores sampling, 1 thread meshing, like the reference): %.3g voxels/s for the step
(sampling alone %.2g voxels/s, meshing %.2g tris/s) → the 1-GPU step is ≈ %s× the CPU step, the e2e call ≈ %s×.

Stage times at 1024³ on one GPU (ms): K1 sample %.2f · K2' classify (sign blocks) %.2f · K3 scan 2 × %.2f · K4a compact %.2f · K4b emit %.2f
(triangle kernel 0.15 + vertex kernel 0.44). ncu launch list of the same command: `profiles/%s_launches_step_summary.txt` (K1 72 %%, emit
16.4 %%, compact 5.4 %%, classify 3.7 %%, scans 2 %% — the same shares as the event-timed stages).

| kernel | algorithmic bytes | time | achieved | fraction of measured HBM peak |
|---|---|---|---|---|
| K1 `sdfk_k_sample` | 16 B × 1.07e9 voxels = 17.18 GB written (ncu: 17.31 GB DRAM writes incl. 134 MB of sign blocks, 6 MB reads) | %.2f ms | %.2f TB/s | **%.2f** |
| K1d `sdfk_k_sample_dist` | 4 B × 1.07e9 = 4.29 GB written (ncu: 4.37 GB) | 0.90 ms | 4.8 TB/s | 0.73 (instruction-issue bound: 83 %% issue-active) |
| K2' `mc_classify_signs` | 1 bit × 1.07e9 = 134 MB read (ncu: 146 MB) | 0.14 ms | — | replaces K2's 4.29 GB pass (0.86 ms, 0.76 of peak) |
| K4b `mc_emit_verts` | 36 B × 3.9e6 vertices written + ≤ 4 × 8 corner reads (ncu: 1.0 GB read, 0.2 GB written) | 0.44 ms | — | latency-bound (≈ 32 %% issue-active, 28 of 32 lanes) |

History of the round (1024³, one GPU, ms/step): first correct path 16.4 → z-column sampling + batched MC loads 7.5 → emit
unrolled / 32-byte records 6.7 → classify counts only active cells 5.8 → lane-per-active-cell compact 5.3 → 5.2 → sign blocks
(K2 0.86 → K2' 0.14) 4.27 → emit split into triangle + kind-sorted vertex kernels, table-driven gathers (0.93 → 0.63), compact at
4 CTAs/SM, second scan over active chunks only (0.175 → 0.087) → %.2f.
e2e `Sdf.ToMesh`: 64 (pageable host buffers) → 9.6 (pinned) → 7.6 (distance-only voxels) → 6.0–6.5 (z-slab pipeline, chunked emit,
streamed downloads; 4.3 ms of it is the 234 MB over PCIe).

## The five BASELINE.json configurations (one GPU; `profiles/%s_configs.txt`)

| config | GPU | CPU restatement (bounded sample, 16 cores) |
|---|---|---|
""" % (multi[8]["value"] / d["value"] if multi[8] else 0, "`profiles/%s_bench_8gpu.json`" % tag, cb["value"], cb["detail"]["sample_voxels_per_s"],
       cb["detail"]["mesh_tris_per_s"], "{:,.0f}".format(d["value"] / cb["value"]), "{:,.0f}".format(d["e2e"]["value"] / cb["value"]),
       st["sample_ms"], st["classify_ms"], st["scan_ms"] / 2, st["compact_ms"], st["emit_ms"], late,
       st["sample_ms"], d5["roofline"]["achieved"] / 1e3, d5["roofline"]["frac"], d5["ms_per_step"], late)
for x in cfg:
    if "render_ms" in x:
        txt += "| %s | %.3f ms = %.3g pixels/s = %.3g SDF evals/s = %.3g FP32 op/s (%.2f of the no-FMA FP32 rate) | %d×%d: %.3g pixels/s |\n" % (
            x["config"], x["render_ms"], x["pixels_per_s"], x["sdf_evals_per_s"], x["fp32_ops_per_s"], x["fp32_ops_per_s"] / FP32,
            x["cpu_image"][0], x["cpu_image"][1], x["cpu_pixels_per_s"])
    else:
        extra = "; %.2f of the no-FMA FP32 rate" % (x["samples_per_s"] * x["sdf_flops_per_sample"] / FP32) if x["sdf_flops_per_sample"] > 100 else ""
        txt += "| %s | sample %.2f ms = %.3g voxels/s (%.2f of HBM peak%s); mesh %.2f ms = %.3g tris/s, %.3g cells/s; %s triangles | %d³: sample %.3g voxels/s; mesh %.3g tris/s |\n" % (
            x["config"], x["sample_ms"], x["samples_per_s"], x["sample_frac_hbm"], extra, x["mesh_ms"], x["tris_per_s"], x["cells_per_s"],
            "{:,}".format(x["triangles"]), x["cpu_grid"], x["cpu_samples_per_s"], x["cpu_tris_per_s"])
if multi[8]:
    txt += "| 4: README scene 2048³ over 8 GPUs | %.2f ms/step, %.3g voxels/s, %.3g tris/s (`profiles/%s_bench_8gpu.json`) | no CPU counterpart (int32 / array limits of the reference) |\n" % (
        multi[8]["ms_per_step"], multi[8]["value"], multi[8]["tris_per_s"], tag)
open(os.path.join(ROOT, "RESULTS.md"), "w").write(txt)
print(txt)
