#!/bin/bash
out=gpurun_out; tag=${TAG:-r03n}
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 600 $out/${tag}_bench_1gpu.err
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_1gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline","fused") if k in d})
print({k:v for k,v in d.items() if k.endswith("_ms") or k in ("stages","mesh")})
PY
