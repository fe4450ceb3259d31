#!/bin/bash
# round-2 GPU pass M: packed masked-chunk path of the tcgen05 attention, tail kernel back to ascending batch order, graph test
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_graph_gpu.py -m gpu -k "attention or graph" --timeout 120 -x -q > $O/pytest_attention_graph.log 2>&1; rc=$?; tail -3 $O/pytest_attention_graph.log
if [ $rc -ne 0 ]; then echo "tests failed (rc=$rc): stopping"; grep -E "timeout|Error|error|assert" $O/pytest_attention_graph.log | head -20; exit 1; fi
for m in 1 0; do echo "L=263 AB_MASKED=$m"; AB_MASKED=$m AB_SPLIT=1 timeout 120 python tools/attn_bench.py; done > $O/attn_bench_263.txt 2>&1; cat $O/attn_bench_263.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-300 $O/bench_cfg3.json
timeout 300 python bench.py --ragged --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-incremental > $O/bench_cfg3_ragged.json 2> $O/bench_cfg3_ragged.err; echo "ragged rc=$?"; cut -c1-300 $O/bench_cfg3_ragged.json
timeout 900 python -m pytest tests -m gpu --timeout 300 -q --deselect tests/test_kernels_gpu.py > $O/pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -5 $O/pytest_rest.log | cut -c1-300
ls -la $O
