#!/bin/bash
# gpurun_out/<tag>_* (tools/measure_round.sh) -> profiles/<tag>_*: bench lines, ncu key metrics per kernel, the launch list's
# per-kernel shares, K1's DRAM traffic.
tag=${1:-r02}
in=gpurun_out; out=profiles
for f in bench_1gpu bench_reference bench_2gpu bench_4gpu bench_8gpu; do [ -s $in/${tag}_$f.json ] && cp $in/${tag}_$f.json $out/; done
for f in kernels tomesh; do [ -s $in/${tag}_$f.txt ] && cp $in/${tag}_$f.txt $out/; done
for k in k1_readme k1_csg50 k1d k2s k4a_compact k4b_emit_tris k4b_emit_verts k5_readme k5_perf; do
  if [ -s $in/${tag}_$k.ncu-rep ]; then
    { echo "# ncu --set full --clock-control none, one launch ($in/${tag}_$k.ncu-rep): key metrics, top stall reasons, opcode histogram";
      python tools/ncu_key.py $in/${tag}_$k.ncu-rep;
      ncu -i $in/${tag}_$k.ncu-rep --page source --csv 2>/dev/null > /tmp/_src.csv; python tools/ncu_src.py /tmp/_src.csv 8; } > $out/${tag}_ncu_$k.txt 2>/dev/null
  fi
done
[ -s $in/${tag}_launches_step.csv ] && python - "$in/${tag}_launches_step.csv" > $out/${tag}_launches_step_summary.txt <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) == len(h):
        try: v = float(r[iv].replace(",", ""))
        except ValueError: continue
        tot[r[ik][:48]] += v; cnt[r[ik][:48]] += 1
s = sum(tot.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none: bench.py --steps 2 --warmup 1 (3 warm-up + 2 timed steps; cold-cache, serialised launches: shares, not absolutes)")
setup = {k: v for k, v in tot.items() if "constdiv_verify" in k or "selftest" in k or "store_probe" in k}
s -= sum(setup.values())
for k, v in tot.most_common():
    if k in setup:
        why = "after the steps: the store-only bandwidth probe behind roofline.write_only_peak" if "store_probe" in k else "at ToSdf() time: exhaustive check of a constant division, not part of a step"
        print("%-50s %4d launches  %10.3f ms total  (%s)" % (k, cnt[k], v / 1e6, why))
    else:
        print("%-50s %4d launches  %10.3f ms total  %5.1f %% of the steps" % (k, cnt[k], v / 1e6, 100 * v / s))
PY
