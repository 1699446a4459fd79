#!/bin/bash
out=gpurun_out; tag=r02o
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import sdfkit_b200 as sk
from sdfkit_b200 import _native as N
n = 512
vals = N.PinnedPool.empty((n, n, n), np.float32); vals[:] = 1.0
cols = N.PinnedPool.empty((n, n, n, 3), np.float32); cols[:] = 0.5
for rep in range(3):
    t0 = time.perf_counter(); v = sk.Voxels(vals, cols, (-1,)*3, (1,)*3); dt = time.perf_counter() - t0
    print("import 512^3 (2.1 GB, pinned source): %.1f ms = %.1f GB/s" % (dt * 1e3, 16.0 * n**3 / dt / 1e9)); v.Dispose()
PY
