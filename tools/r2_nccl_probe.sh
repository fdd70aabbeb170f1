#!/bin/bash
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=29550
probe() { tag=$1; shift; p=$((p+1)); env "$@" $RUN --master-port $p tools/nccl_allreduce_probe.py $tag 2>/dev/null | grep all-reduce >> gpurun_out/r2_nccl_probe_${N}gpu.log; }
: > gpurun_out/r2_nccl_probe_${N}gpu.log
probe default NCCL_DEBUG=WARN
probe algo_nvls NCCL_ALGO=NVLS
probe algo_ring NCCL_ALGO=Ring
probe algo_tree NCCL_ALGO=Tree
probe nch32 NCCL_MIN_NCHANNELS=32
probe ctas32 NCCL_MIN_CTAS=32 NCCL_MAX_CTAS=64
probe nvls_off NCCL_NVLS_ENABLE=0
cat gpurun_out/r2_nccl_probe_${N}gpu.log
