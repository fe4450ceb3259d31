#!/bin/bash
# round-2 GPU pass N: final tree: full GPU suite (incl. the non-GEGLU fixture test), default bench
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu --timeout 300 -q > $O/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 $O/pytest_all.log | cut -c1-300; grep -E "^FAILED|Error" $O/pytest_all.log | head
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"; cut -c1-300 $O/bench_default.json; tail -4 $O/bench_default.err
ls -la $O
