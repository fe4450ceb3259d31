#!/bin/bash
# round-2 GPU pass E: precision diagnosis per stage + ncu --set full captures of the stats GEMM, the GEGLU GEMM and the attention kernel
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python tools/precision_diag.py > $O/precision_diag.txt 2>&1; cat $O/precision_diag.txt
timeout 900 ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "timed/" --kernel-name-base demangled \
  --kernel-name 'regex:attention_tc_kernel|EpiCfg<0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1>|EpiCfg<0, 3, 1, 0, 0, 0, 1, 0, 1, 0, 0>|EpiCfg<0, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0>' \
  --launch-skip 8 --launch-count 8 -f -o $O/full_layer0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/ncu_full.log
ls -la $O
