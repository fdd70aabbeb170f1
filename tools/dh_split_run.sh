#!/bin/bash
# One GPU call: correctness of the dHidden split-K path, interleaved A/B timing, bench at the reference's micro-batch
# size with and without it, then the whole GPU suite with the option on. Logs under gpurun_out/dh_split/.
out=gpurun_out/dh_split
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt 2>&1
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "multi_round_split or dhidden_split or split_k_tail" > $out/tests_new.log 2>&1
echo "new tests rc=$?" | tee -a $out/summary.txt
for rows in 4096 1024 8192; do
  timeout 120 python tools/gpu_ab.py "dh_split=0" "dh_split=1" --rows $rows --rounds 4 --iters 10 > $out/ab_rows$rows.log 2>&1
  echo "ab rows=$rows rc=$?" | tee -a $out/summary.txt
done
timeout 120 python tools/gpu_ab.py "dh_split=0" "dh_split=1" --rows 4096 --hidden 2048 --rounds 4 --iters 10 > $out/ab_rows4096_h2048.log 2>&1
for s in 0 1; do
  GRPO_DH_SPLIT=$s timeout 200 python bench.py --sequences 512 --micro-seqs 4 --steps 2 --warmup 3 --no-e2e --no-cpu > $out/bench_m4_split$s.json 2> $out/bench_m4_split$s.err
  echo "bench split=$s rc=$?" | tee -a $out/summary.txt
done
GRPO_DH_SPLIT=1 timeout 600 python -m pytest tests -m gpu -x -q > $out/tests_all_split1.log 2>&1
echo "full suite (dh_split=1) rc=$?" | tee -a $out/summary.txt
tail -3 $out/tests_all_split1.log
cat $out/summary.txt
