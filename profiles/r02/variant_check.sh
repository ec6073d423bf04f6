#!/bin/bash
# usage: bash profiles/r02/variant_check.sh NAME...   (variants built by profiles/build_variant.sh)
# per variant: randomised differential run (two seeds), common-path timing (N(0,1)), long-run timing (check_runs.py)
mkdir -p gpurun_out
for v in "$@"; do
  export SPECKV_LIB=cxl_speckv_b200/build/variants/lib$v.so
  [ "$v" = "tree" ] && unset SPECKV_LIB
  echo "=== $v"
  for seed in 12 13; do
    timeout 200 python tests/fuzz_codec.py 40 $seed 2>&1 | grep -v "RuntimeWarning\|x \*=" | tail -2
  done
  timeout 120 python profiles/quick_time.py 131072 8192 20 2>&1 | tail -1
  timeout 200 python profiles/r02/check_runs.py 2>&1 | grep "n= 512" | grep "scheme 2"
done
