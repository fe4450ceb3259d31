#!/bin/bash
# round-2 GPU pass K: the driver's own commands on the current tree: default bench (both arms), smoke, full GPU suite
O=gpurun_out/r2k; mkdir -p $O
AB_SPLIT=1 timeout 120 python tools/attn_bench.py > $O/attn_bench_263.txt 2>&1; cat $O/attn_bench_263.txt
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"; cut -c1-300 $O/bench_default.json; tail -4 $O/bench_default.err
( time timeout 600 python bench.py --impl reference ) > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; cut -c1-400 $O/bench_reference.json; tail -4 $O/bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu --timeout 300 -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 $O/pytest_all.log | cut -c1-300
ls -la $O
