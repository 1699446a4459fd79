#!/bin/bash
# compute-sanitizer over the round-2 code paths (multi-GPU layer, shared-guard device forms, 8-per-lane sampler, neighbourhood
# emit, chunked export, pipelined Sdf.ToMesh on two streams)
out=gpurun_out; tag=r02_sanitizer
export SDFK_RENDER_BANDS=3   # small test images also take the banded render -> copy pipeline
{
echo "# compute-sanitizer on the B200 box (round 2)"
K1="multi_to_mesh and readme-dims0 or sub_slabs and 5 or sharded_voxels and perf or multi_render or chunked_voxel_export and dims1 or depth_tga or indexer"
echo "## memcheck: pytest tests/test_gpu_multi.py -k '$K1'"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py -x -q -k "$K1" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -20
K2="store_bandwidth or render_tga or fused and 256 or 8_per_lane or default_device_forms and readme or pipelined and dims1 or white_noise_all and 24 or mesh_matches and perf or step_and_iso"
echo "## memcheck: pytest tests/test_gpu_parity.py -k '$K2'"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$K2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -20
K3="multi_to_mesh and sphere-dims1 or 8_per_lane and dims2"
echo "## initcheck: -k '$K3'"
compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "$K3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Uninit|Error" | head -20
K4="white_noise_all and 24 or multi_to_mesh and perf"
echo "## racecheck: -k '$K4'"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "$K4" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" | head -20
} > $out/${tag}.txt 2>&1
cat $out/${tag}.txt
