#!/bin/bash
# round-2 GPU pass F: ncu --set full of every decoder GEMM shape in f16f8 (microbench), attention tail-tile experiment (L = 256 vs 263)
O=gpurun_out/r2f; mkdir -p $O
for L in 263 256 264 384 392; do echo "L=$L"; AB_L=$L AB_SPLIT=1 timeout 100 python tools/attn_bench.py; done > $O/attn_tail.txt 2>&1; cat $O/attn_tail.txt
GB_SPLIT=2 GB_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 18 -f -o $O/gemm_f16f8_full python tools/gemm_bench.py > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; grep "split=" $O/ncu_gemm.log
ls -la $O
