#!/bin/bash
# same-box A/B: previous build vs this build (6e6 candidates, N_train=2000, d=12)
for rep in 1 2; do
for v in prev new; do
  lib=gpry_b200/libgpry_b200.so; [ "$v" = prev ] && lib=gpry_b200/libgpry_b200_prev.so
  echo -n "$v: "
  GPRY_B200_LIB=$PWD/$lib python tools/probe_contract.py 6000000 2000 12 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('contract_ms %.1f build %.1f total %.1f' % (d['stage_ms']['contract_ms'], d['stage_ms']['build_ms'], d['ms_per_pass']))"
  GPRY_B200_LIB=$PWD/$lib python -c "
from gpry_b200 import DeviceGP
d=DeviceGP(0); print('   int8 peak TOPS burst %.0f sustained %.0f' % (d.int8_peak_tops(), d.int8_peak_tops(seconds=1.0)))"
done
done
