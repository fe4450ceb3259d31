#!/bin/bash
# round-2 GPU pass G: SIMT tail kernel for the rows past the last full 128-row attention tile; CUDA-graph replay as the bench default
O=gpurun_out/r2g; mkdir -p $O
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -k "attention" --timeout 90 -x -q > $O/pytest_attention.log 2>&1; rc=$?; tail -5 $O/pytest_attention.log
if [ $rc -ne 0 ]; then echo "attention tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_attention.log | head -20; exit 1; fi
AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_tail1.txt 2>&1; cat $O/attn_bench_tail1.txt
VIMA_B200_ATTN_TAIL=0 AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_tail0.txt 2>&1; cat $O/attn_bench_tail0.txt
AB_L=392 AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_392_tail1.txt 2>&1; cat $O/attn_bench_392_tail1.txt
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-graph > $O/bench_cfg3_nograph.json 2> $O/bench_cfg3_nograph.err; echo "cfg3 nograph rc=$?"; cut -c1-300 $O/bench_cfg3_nograph.json
timeout 1200 python -m pytest tests -m gpu --timeout 300 -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/pytest_all.log | cut -c1-400
timeout 300 python bench.py --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 $O/bench_cfg5.json
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-300 $O/bench_cfg2.json
ls -la $O
