#!/bin/bash
# Deferred dW (fused.DeferredDW): its two GPU tests, the whole GPU suite, then bench at the reference's 4-sequence
# micro-batches without / with --defer-dw on the same box.
out=gpurun_out/defer
mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "deferred_dw" > $out/tests_new.log 2>&1
echo "new tests rc=$?"; tail -n 4 $out/tests_new.log
timeout 300 python -m pytest tests -m gpu -x -q > $out/tests_all.log 2>&1
echo "full suite rc=$?"; tail -n 2 $out/tests_all.log
for f in "" "--defer-dw"; do
  timeout 100 python bench.py --sequences 256 --micro-seqs 4 --steps 2 --warmup 3 --no-e2e --no-cpu $f > $out/bench_m4_defer${f:+1}.json 2> $out/bench_m4_defer${f:+1}.err
  echo "bench '$f' rc=$?"; cut -c1-160 $out/bench_m4_defer${f:+1}.json
done
