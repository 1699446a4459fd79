#!/bin/bash
out=gpurun_out
K2="fused and 256 or 8_per_lane or default_device_forms and readme or pipelined and dims1 or white_noise_all and 24 or mesh_matches and perf or step_and_iso"
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "$K2" > $out/r02_san_parity_full.txt 2>&1
K1="multi_to_mesh and readme-dims0 or sub_slabs and 5 or sharded_voxels and perf or multi_render or chunked_voxel_export and dims1 or depth_tga or indexer"
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_multi.py -x -q -k "$K1" > $out/r02_san_multi_full.txt 2>&1
grep -c "=========" $out/r02_san_parity_full.txt $out/r02_san_multi_full.txt
