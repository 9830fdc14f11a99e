#!/bin/bash
# occupancy / slicing experiments on kstar_build (INT8 mode, N_train=2000, d=12, 6e6 candidates)
for cfg in "0 8" "46 8" "57 8" "0 16" "29 16" "33 16" "0 4"; do
  set -- $cfg
  echo -n "smem_kb=$1 JS=$2: "
  GPRY_B200_BUILD_SMEM_KB=$1 GPRY_B200_BUILD_JS=$2 python tools/probe_contract.py 6000000 2000 12 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('build_ms %.1f contract_ms %.1f total %.1f' % (d['stage_ms']['build_ms'], d['stage_ms']['contract_ms'], d['ms_per_pass']))"
done
