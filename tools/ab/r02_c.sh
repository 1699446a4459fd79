#!/bin/bash
# 2-GPU pass: multi-GPU tests on real devices (incl. the torchrun 2-rank job) and the N=2 bench line
out=gpurun_out; tag=r02c
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_pytest_multi.log 2>&1; tail -3 $out/${tag}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err; tail -c 1500 $out/${tag}_bench_2gpu.err
