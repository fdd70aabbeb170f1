#!/bin/bash
# Config C5 (ragged 4096-token responses, padded rows compacted away -> chunk sizes that are not whole rounds of tiles):
# bench line with the dHidden split-K path off and on, same box, back to back.
out=gpurun_out/final
mkdir -p $out
for s in 0 1; do
  GRPO_DH_SPLIT=$s timeout 110 python bench.py --config c5 --sequences 512 --steps 2 --warmup 3 --no-e2e --no-cpu > $out/bench_c5_split$s.json 2> $out/bench_c5_split$s.err
  echo "c5 split=$s rc=$?"
  cut -c1-200 $out/bench_c5_split$s.json
done
