M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
run() { tag=$1; shift; env "$@" ncu --metrics $M --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv --log-file gpurun_out/knob3_$tag.csv python tools/gpu_prof_target.py 3584 9472 2 > /dev/null 2>&1; }
run base GRPO_SYNC_DH=0
run s16 GRPO_SYNC_DH=16 GRPO_SYNC_DW=16 GRPO_SYNC_FWD=56
run s8 GRPO_SYNC_DH=8 GRPO_SYNC_DW=8 GRPO_SYNC_FWD=28
run s32 GRPO_SYNC_DH=32 GRPO_SYNC_DW=37 GRPO_SYNC_FWD=112
