#!/bin/bash
# Round 2 ncu evidence: (1) launch list of one bench step (1/8 of the C3 batch = what each rank of the 8-GPU run executes),
# (2) --set full capture of the three GEMMs of one 18 944-row chunk (logits+stash, dHidden, dW).
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launch_list_raw.csv \
    python bench.py --sequences 512 --steps 1 --warmup 0 --no-e2e --no-cpu --no-records > gpurun_out/r2_launch_list_bench.json 2> gpurun_out/r2_launch_list_bench.err
tail -2 gpurun_out/r2_launch_list_bench.err; wc -l gpurun_out/r2_launch_list_raw.csv
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 3 -f -o gpurun_out/r2_prof_gemms \
    python tools/gpu_prof_target.py 3584 18944 2 > /dev/null 2>&1
ls -la gpurun_out/r2_prof_gemms.ncu-rep
