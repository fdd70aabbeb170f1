#!/bin/bash
# Round 2: per-operand L2 eviction priorities on the three GEMMs (option l2_hints bits 4 / 8 / 16): DRAM bytes under ncu and
# interleaved wall-clock A/B.   gpurun -- bash tools/r2_l2_hints.sh
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
run() { tag=$1; shift; env "$@" ncu --metrics $M --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv --log-file gpurun_out/r2_l2_$tag.csv python tools/gpu_prof_target.py 3584 18944 2 > /dev/null 2>&1; }
run base GRPO_L2_HINTS=0
run dw_hd_last GRPO_L2_HINTS=4
run fwd_hid_last GRPO_L2_HINTS=8
run dh_w_last GRPO_L2_HINTS=16
python tools/ncu_summary.py gpurun_out/r2_l2_*.csv > gpurun_out/r2_l2_hints_ncu.txt 2>&1
cat gpurun_out/r2_l2_hints_ncu.txt
python tools/gpu_ab.py "l2_hints=0" "l2_hints=4" "l2_hints=8" "l2_hints=16" "l2_hints=28" --rows 18944 --rounds 5 --iters 6 > gpurun_out/r2_ab_l2_hints.log 2>&1
tail -30 gpurun_out/r2_ab_l2_hints.log
