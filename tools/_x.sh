compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined and (dims1 or dims4 or dims6) or chunked_emit or white_noise_all" > gpurun_out/s2_memcheck.log 2>&1
grep -m3 -A12 "Invalid\|Error\|error:" gpurun_out/s2_memcheck.log | head -60
grep -c "Invalid" gpurun_out/s2_memcheck.log
