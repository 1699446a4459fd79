#!/bin/bash
out=gpurun_out; tag=r02e
python -m pytest tests/test_gpu_parity.py -x -q -k "default_device_forms or mesh_matches or sample_matches or render_matches" > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
for w in 0 2 6 100; do echo "WIDE_MAX=$w"; SDFK_WIDE_MAX=$w python tools/time_kernels.py; done > $out/${tag}_kernels.txt 2>&1
cat $out/${tag}_kernels.txt
