#!/bin/bash
# usage: run_variants.sh "<tune_chord args>" name1 name2 ...   -- times scripts/tune_chord.py under each csrc/build/variants/<name>.so
ARGS="$1"; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== $v"
  LBM_B200_LIB=$PWD/pour_over_coffee_lbm_b200/csrc/build/variants/$v.so python scripts/tune_chord.py $ARGS 2>&1 | grep -v "^$" | sed "s/^/[$v] /"
done
