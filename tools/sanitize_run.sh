#!/bin/bash
# compute-sanitizer memcheck over the GPU tests of the split-K paths and the fused-loss tests (default options).
out=gpurun_out/final
mkdir -p $out
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "multi_round_split or dhidden_split or split_k_tail or fused_grpo_loss or edge_cases" > $out/memcheck_split.log 2>&1
echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY" $out/memcheck_split.log
