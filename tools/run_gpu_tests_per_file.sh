#!/bin/bash
# every GPU test file in its own process (a CUDA fault in one file does not poison the others); summary lines only
mkdir -p gpurun_out
out=${1:-gpurun_out/per_file.log}
: > $out
for f in tests/test_gpu_*.py tests/test_match_golden.py; do
  echo "== $f" >> $out
  timeout 600 python -m pytest $f -m gpu -q -x 2>&1 | tail -${2:-15} >> $out
done
grep -E "^==|passed|failed|error" $out
