#!/bin/bash
# round-2 GPU pass I: tail rows fused into the tcgen05 attention kernel (vs launch of their own vs one more tile)
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -k attention --timeout 120 -x -q > $O/pytest_attention.log 2>&1; rc=$?; tail -5 $O/pytest_attention.log
if [ $rc -ne 0 ]; then echo "attention tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_attention.log | head -20; exit 1; fi
for mode in fused kernel off; do echo "L=263 attn_tail=$mode"; VIMA_B200_ATTN_TAIL=$mode AB_SPLIT=1 timeout 120 python tools/attn_bench.py; done > $O/attn_bench_263.txt 2>&1; cat $O/attn_bench_263.txt
for mode in fused off; do echo "L=392 attn_tail=$mode"; AB_L=392 VIMA_B200_ATTN_TAIL=$mode AB_SPLIT=1 timeout 120 python tools/attn_bench.py; done > $O/attn_bench_392.txt 2>&1; cat $O/attn_bench_392.txt
timeout 400 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-400 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
timeout 1200 python -m pytest tests -m gpu --timeout 300 -q --deselect tests/test_kernels_gpu.py > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -8 $O/pytest_rest.log | cut -c1-400
timeout 300 python bench.py --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 $O/bench_cfg5.json
timeout 300 python bench.py --workload cfg3x --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/bench_cfg3x.json 2> $O/bench_cfg3x.err; echo "cfg3x rc=$?"; cut -c1-300 $O/bench_cfg3x.json
ls -la $O
