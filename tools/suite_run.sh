#!/bin/bash
out=gpurun_out/defer
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/tests_all_defer_default.log 2>&1
echo "full suite rc=$?"; tail -n 2 $out/tests_all_defer_default.log
