#!/bin/bash
# round-2 GPU pass B: new attention kernel + folded-LN GEMMs: kernel tests, policy tests, microbenches, step bench, ncu step profile
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_cabi_exports.py -m gpu -x -q > $O/pytest_kernels.log 2>&1; echo "kernels rc=$?"; tail -4 $O/pytest_kernels.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_kernels_gpu.py > $O/pytest_rest.log 2>&1; echo "rest rc=$?"; tail -4 $O/pytest_rest.log
AB_B=256 timeout 200 python tools/attn_bench.py > $O/attn_bench_tc.txt 2>&1; cat $O/attn_bench_tc.txt
VIMA_B200_ATTN=mma AB_B=256 timeout 200 python tools/attn_bench.py > $O/attn_bench_mma.txt 2>&1
GB_REPS=5 timeout 300 python tools/gemm_bench.py > $O/gemm_bench.txt 2>&1; grep "split=2" $O/gemm_bench.txt
timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 $O/bench_cfg3.json
timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg3_graph.json 2> $O/bench_cfg3_graph.err; echo "cfg3 graph rc=$?"; cut -c1-300 $O/bench_cfg3_graph.json
timeout 600 python bench.py --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 $O/bench_cfg5.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/kernel_metrics_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
ls -la $O
# cfg3x parity triangulation (ours vs the eager reference on the same GPU at Lp=512)
timeout 500 python bench.py --workload cfg3x --steps 3 --warmup 3 --no-cpu-baseline --no-incremental > $O/bench_cfg3x_f16f8.json 2> $O/bench_cfg3x_f16f8.err
timeout 500 python bench.py --workload cfg3x --steps 3 --warmup 3 --no-cpu-baseline --no-incremental --precision f16x3 > $O/bench_cfg3x_f16x3.json 2> $O/bench_cfg3x_f16x3.err
VIMA_B200_ATTN=mma timeout 500 python bench.py --workload cfg3x --steps 3 --warmup 3 --no-cpu-baseline --no-incremental --precision f16x3 > $O/bench_cfg3x_f16x3_mma.json 2> $O/bench_cfg3x_f16x3_mma.err
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg2_graph.json 2> $O/bench_cfg2_graph.err; echo "cfg2 graph rc=$?"; cut -c1-300 $O/bench_cfg2_graph.json
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-300 $O/bench_cfg2.json
ls -la $O
