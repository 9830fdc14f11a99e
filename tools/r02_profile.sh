#!/bin/bash
# round-2 evidence: launch list of the bench command, ncu --set full of the two dominant kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oz2_contract -s 1 -c 1 -o gpurun_out/r02_oz2 \
    python tools/probe_contract.py 80000 2000 12 > gpurun_out/r02_ncu_oz2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kstar_build -s 2 -c 1 -o gpurun_out/r02_build \
    python tools/probe_contract.py 80000 2000 12 > gpurun_out/r02_ncu_build.log 2>&1
ncu --set full --clock-control none -k regex:gemm_nt -s 150 -c 3 -o gpurun_out/r02_gemm \
    python tools/probe_train.py 4000 20 8 > gpurun_out/r02_ncu_gemm.log 2>&1
