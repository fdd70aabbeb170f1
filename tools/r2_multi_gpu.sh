#!/bin/bash
# Round 2: configs C3 / C4 / C5 at their stated size on N GPUs of one box (gpurun --gpus N -- bash tools/r2_multi_gpu.sh N)
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8 > gpurun_out/r2_gpus_${N}.txt
free -g | head -2 >> gpurun_out/r2_gpus_${N}.txt
$RUN --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_c3_${N}gpu.json 2> gpurun_out/r2_bench_c3_${N}gpu.err
cut -c1-600 gpurun_out/r2_bench_c3_${N}gpu.json; tail -2 gpurun_out/r2_bench_c3_${N}gpu.err
$RUN --master-port 29512 bench.py --gpus $N --config c4 --steps 4 --warmup 2 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
cut -c1-600 gpurun_out/r2_bench_c4_${N}gpu.json; tail -2 gpurun_out/r2_bench_c4_${N}gpu.err
$RUN --master-port 29513 bench.py --gpus $N --config c5 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
cut -c1-600 gpurun_out/r2_bench_c5_${N}gpu.json; tail -2 gpurun_out/r2_bench_c5_${N}gpu.err
$RUN --master-port 29514 bench.py --gpus $N --config c5 --steps 3 --warmup 2 --no-e2e --no-cpu --no-records --balance 0 > gpurun_out/r2_bench_c5_${N}gpu_unbalanced.json 2> gpurun_out/r2_bench_c5_${N}gpu_unbalanced.err
cut -c1-400 gpurun_out/r2_bench_c5_${N}gpu_unbalanced.json; tail -2 gpurun_out/r2_bench_c5_${N}gpu_unbalanced.err
