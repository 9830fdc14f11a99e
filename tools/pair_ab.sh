#!/bin/bash
# same-box A/B: previous build (one digit per MMA, [digit][k16] V layout) vs paired N = 256 MMAs
for rep in 1 2; do
for v in "prev 0" "new 0" "new 32"; do
  set -- $v
  lib=gpry_b200/libgpry_b200.so; [ "$1" = prev ] && lib=gpry_b200/libgpry_b200_prev.so
  echo -n "$1 dbg=$2: "
  if [ "$2" = 0 ]; then unset GPRY_B200_OZ_DBG; else export GPRY_B200_OZ_DBG=$2; fi
  GPRY_B200_LIB=$PWD/$lib python tools/probe_contract.py 6000000 2000 12 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('contract_ms %.1f build %.1f total %.1f' % (d['stage_ms']['contract_ms'], d['stage_ms']['build_ms'], d['ms_per_pass']))"
done
done
