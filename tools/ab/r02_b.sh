#!/bin/bash
out=gpurun_out; tag=r02b
python tools/make_mesh_digests.py $out/${tag}_mesh_digests.json > $out/${tag}_digests.log 2>&1
cp $out/${tag}_mesh_digests.json tests/golden/mesh_digests.json
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 600 $out/${tag}_bench_1gpu.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
