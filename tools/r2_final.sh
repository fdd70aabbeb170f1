#!/bin/bash
# Round-2 closing run on one B200: the whole GPU suite, smoke, the default bench line, memcheck of the peer kernels.
#   gpurun -- bash tools/r2_final.sh
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_final_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | tee -a gpurun_out/r2_final_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2_bench_c3_1gpu_final.json 2> gpurun_out/r2_bench_c3_1gpu_final.err
cut -c1-400 gpurun_out/r2_bench_c3_1gpu_final.json; tail -2 gpurun_out/r2_bench_c3_1gpu_final.err
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_peer.py -m gpu -q -k "local_buffers and 32792" 2>&1 | tail -6 | tee gpurun_out/r2_sanitizer_peer.txt
