#!/bin/bash
# A/B of the split-K GEMM on the training side (LML+grad per evaluation, factorisation latency)
for sk in 0 1; do
  echo "GPRY_B200_SPLITK=$sk"
  for cfg in "2000 12 1" "2000 12 8" "4000 20 1" "4000 20 8" "4000 20 16" "4000 20 64"; do
    GPRY_B200_SPLITK=$sk python tools/probe_train.py $cfg 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('  ', '$cfg', 'lml_grad_ms_per_eval %.3f factorize_ms %.2f' % (d['lml_grad_ms_per_eval'], d['factorize_device_ms']), d.get('lml_relerr'), d.get('grad_err'))"
  done
done
