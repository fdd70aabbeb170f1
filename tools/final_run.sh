#!/bin/bash
# Round-end evidence in one GPU call: ncu of the dHidden GEMM at the reference's micro-batch size with and without the
# split-K path, then the default bench line and the reference arm.
out=gpurun_out/final
mkdir -p $out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
GRPO_DH_SPLIT=0 timeout 200 ncu --metrics $M --clock-control none -k 'regex:gemm_kernel|dh_fixup' -s 4 -c 4 --csv --log-file $out/ncu_dh_rows4096_split0.csv python tools/gpu_prof_target.py 3584 4096 2 > /dev/null 2>&1
GRPO_DH_SPLIT=1 timeout 200 ncu --metrics $M --clock-control none -k 'regex:gemm_kernel|dh_fixup' -s 5 -c 5 --csv --log-file $out/ncu_dh_rows4096_split1.csv python tools/gpu_prof_target.py 3584 4096 2 > /dev/null 2>&1
python tools/ncu_summary.py $out/ncu_dh_rows4096_split0.csv $out/ncu_dh_rows4096_split1.csv > $out/ncu_dh_rows4096.txt 2>&1
timeout 200 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
echo "reference arm rc=$?"
timeout 400 python bench.py > $out/bench_c3_1gpu.json 2> $out/bench_c3_1gpu.err
echo "bench rc=$?"
cat $out/ncu_dh_rows4096.txt
cut -c1-300 $out/bench_c3_1gpu.json
