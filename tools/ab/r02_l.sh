#!/bin/bash
# 4-GPU pass: the N=4 bench line (weak 1624^3, parity, strong_1024 on 1/2/4 devices, config 4 at 2/4)
out=gpurun_out; tag=r02l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > $out/${tag}_bench_4gpu.json 2> $out/${tag}_bench_4gpu.err; tail -c 800 $out/${tag}_bench_4gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-strong > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err; tail -c 300 $out/${tag}_bench_2gpu.err
