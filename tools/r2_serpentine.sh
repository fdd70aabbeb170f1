#!/bin/bash
# Round 2: serpentine K order (option k_serp, bit 0 dW GEMM, bit 1 dHidden GEMM): parity at the headline shape, DRAM bytes
# under ncu, interleaved wall-clock A/B.   gpurun -- bash tools/r2_serpentine.sh
GRPO_K_SERP=3 timeout 200 python -m pytest tests/test_headline_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_serp_parity.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
run() { tag=$1; shift; env "$@" timeout 120 ncu --metrics $M --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv --log-file gpurun_out/r2_serp_$tag.csv python tools/gpu_prof_target.py 3584 18944 2 > /dev/null 2>&1; }
run base GRPO_K_SERP=0
run on GRPO_K_SERP=3
python tools/ncu_summary.py gpurun_out/r2_serp_base.csv gpurun_out/r2_serp_on.csv > gpurun_out/r2_serp_ncu.txt 2>&1
cat gpurun_out/r2_serp_ncu.txt
timeout 120 python tools/gpu_ab.py "k_serp=0" "k_serp=1" "k_serp=3" --rows 18944 --rounds 5 --iters 6 > gpurun_out/r2_ab_serp.log 2>&1
tail -5 gpurun_out/r2_ab_serp.log
