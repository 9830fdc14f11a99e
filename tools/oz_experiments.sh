#!/bin/bash
# timing experiments on the INT8 contraction (GPRY_B200_OZ_DBG bits: 1 no operand traffic,
# 2 MMAs of N=64, 4 epilogue of 16 columns, 8 no V traffic); results are wrong on purpose
for dbg in 16 1 4 8 5 13; do
  echo "dbg=$dbg"
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > /tmp/clk_$dbg.csv &
  SMI=$!
  GPRY_B200_OZ_DBG=$dbg python tools/probe_contract.py 6000000 2000 12 2>&1 | tail -1
  kill $SMI
  sort -t, -k2 -n -r /tmp/clk_$dbg.csv | head -2
done
