#!/bin/bash
# round-2 GPU pass L: current tree: kernel tests (SIMT GEMM prefetch), default bench wall time, ragged line, prompt-encode ncu table, step ncu table
O=gpurun_out/r2l; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_policy_gpu.py tests/test_graph_gpu.py -m gpu --timeout 120 -x -q > $O/pytest_kernels_policy.log 2>&1; rc=$?; tail -3 $O/pytest_kernels_policy.log
if [ $rc -ne 0 ]; then echo "tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_kernels_policy.log | head -20; exit 1; fi
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"; cut -c1-300 $O/bench_default.json; tail -4 $O/bench_default.err
timeout 300 python bench.py --ragged --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/bench_cfg3_ragged.json 2> $O/bench_cfg3_ragged.err; echo "ragged rc=$?"; cut -c1-300 $O/bench_cfg3_ragged.json
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/kernel_metrics_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/ncu_metrics.log 2>&1; echo "ncu step rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "prompt/" --csv --log-file $O/kernel_metrics_prompt.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental --no-graph > $O/ncu_prompt.log 2>&1; echo "ncu prompt rc=$?"
ls -la $O
