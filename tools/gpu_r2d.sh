#!/bin/bash
# round-2 GPU pass D (re-entry): attention v2 first (short fuse), microbench, step bench, then the full suite and the ncu step profile
O=gpurun_out/r2d; mkdir -p $O
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" --timeout 90 -x -q > $O/pytest_attention.log 2>&1; rc=$?; tail -5 $O/pytest_attention.log
if [ $rc -ne 0 ]; then echo "attention tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_attention.log | head -20; exit 1; fi
AB_B=256 timeout 120 python tools/attn_bench.py > $O/attn_bench_tc.txt 2>&1; cat $O/attn_bench_tc.txt
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-600 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg3_graph.json 2> $O/bench_cfg3_graph.err; echo "cfg3 graph rc=$?"; cut -c1-300 $O/bench_cfg3_graph.json; tail -3 $O/bench_cfg3_graph.err
timeout 1200 python -m pytest tests -m gpu --timeout 300 -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/pytest_all.log | cut -c1-400
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/kernel_metrics_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg2_graph.json 2> $O/bench_cfg2_graph.err; echo "cfg2 graph rc=$?"; cut -c1-300 $O/bench_cfg2_graph.json
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-300 $O/bench_cfg2.json
timeout 300 python bench.py --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 $O/bench_cfg5.json
ls -la $O
