#!/bin/bash
# round-2 GPU pass H: pipelined tail kernel + register-prefetched GEMM epilogue residuals
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu --timeout 120 -x -q > $O/pytest_kernels.log 2>&1; rc=$?; tail -5 $O/pytest_kernels.log
if [ $rc -ne 0 ]; then echo "kernel tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_kernels.log | head -20; exit 1; fi
AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_tail1.txt 2>&1; cat $O/attn_bench_tail1.txt
AB_L=392 AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_392_tail1.txt 2>&1; cat $O/attn_bench_392_tail1.txt
GB_SPLIT=2 GB_REPS=5 timeout 200 python tools/gemm_bench.py > $O/gemm_bench_pf0.txt 2>&1; cat $O/gemm_bench_pf0.txt
VIMA_B200_EPI_PREFETCH=1 GB_SPLIT=2 GB_REPS=5 timeout 200 python tools/gemm_bench.py > $O/gemm_bench_pf1.txt 2>&1; cat $O/gemm_bench_pf1.txt
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
VIMA_B200_EPI_PREFETCH=1 timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/bench_cfg3_pf1.json 2> $O/bench_cfg3_pf1.err; echo "cfg3 pf1 rc=$?"; cut -c1-300 $O/bench_cfg3_pf1.json
timeout 1200 python -m pytest tests -m gpu --timeout 300 -q --deselect tests/test_kernels_gpu.py > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -8 $O/pytest_rest.log | cut -c1-400
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/kernel_metrics_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental --no-graph > $O/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
ls -la $O
