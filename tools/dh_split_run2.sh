#!/bin/bash
# Whole GPU suite with the dHidden split-K path on (default) and off, then the smoke entry.
out=gpurun_out/dh_split
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/tests_all_default.log 2>&1
echo "full suite (default, dh_split=1) rc=$?" | tee -a $out/summary2.txt
GRPO_DH_SPLIT=0 timeout 600 python -m pytest tests -m gpu -x -q > $out/tests_all_split0.log 2>&1
echo "full suite (GRPO_DH_SPLIT=0) rc=$?" | tee -a $out/summary2.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a $out/summary2.txt
tail -2 $out/tests_all_default.log $out/tests_all_split0.log $out/smoke.log
