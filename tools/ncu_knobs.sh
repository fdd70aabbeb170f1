# DRAM traffic / tensor-pipe activity of the four GEMM launches of one chunk under knob settings (env GRPO_*).
#   bash tools/ncu_knobs.sh <tag-prefix> <rows>     -> gpurun_out/<tag-prefix>_<variant>.csv
P=${1:-knob}; ROWS=${2:-18944}
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
run() { tag=$1; shift; env "$@" ncu --metrics $M --clock-control none -k regex:gemm_kernel -s 4 -c 4 --csv --log-file gpurun_out/${P}_$tag.csv python tools/gpu_prof_target.py 3584 $ROWS 2 > /dev/null 2>&1; }
run base GRPO_ST_HINT=3
run nohint GRPO_ST_HINT=0
run stcs GRPO_EPI_MODE=1
run panel9472 GRPO_FWD_PANEL=9472
run panel2432 GRPO_FWD_PANEL=2432
run l2h1 GRPO_L2_HINTS=1
run l2h2 GRPO_L2_HINTS=2
