#!/bin/bash
# Round 2, second 8-GPU pass: per-rank all-reduce / kernel times on C3, and C5 with every optimizer step balanced
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$RUN --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_${N}gpu_v2.json 2> gpurun_out/r2_bench_c3_${N}gpu_v2.err
cut -c1-300 gpurun_out/r2_bench_c3_${N}gpu_v2.json; tail -2 gpurun_out/r2_bench_c3_${N}gpu_v2.err
$RUN --master-port 29522 bench.py --gpus $N --config c5 --steps 3 --warmup 2 --no-e2e --no-cpu --no-records > gpurun_out/r2_bench_c5_${N}gpu_v2.json 2> gpurun_out/r2_bench_c5_${N}gpu_v2.err
cut -c1-300 gpurun_out/r2_bench_c5_${N}gpu_v2.json; tail -2 gpurun_out/r2_bench_c5_${N}gpu_v2.err
