#!/bin/bash
# round-2 GPU pass A: tests, the four bench workloads, CUDA-graph A/B on cfg2, CPU thread probe, epilogue-prefetch A/B
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2_f16f8.json 2> $O/bench_cfg2_f16f8.err; echo "cfg2 rc=$?"
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg2_f16f8_graph.json 2> $O/bench_cfg2_f16f8_graph.err; echo "cfg2 graph rc=$?"
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --precision bf16 --no-gpu-eager > $O/bench_cfg2_bf16.json 2> $O/bench_cfg2_bf16.err; echo "cfg2 bf16 rc=$?"
timeout 400 python bench.py --workload cfg2 --steps 20 --warmup 5 --precision bf16 --no-gpu-eager --no-cpu-baseline --graph > $O/bench_cfg2_bf16_graph.json 2> $O/bench_cfg2_bf16_graph.err; echo "cfg2 bf16 graph rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 6 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg3x --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_cfg3x.json 2> $O/bench_cfg3x.err; echo "cfg3x rc=$?"
timeout 400 python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager --graph > $O/bench_cfg3_graph.json 2> $O/bench_cfg3_graph.err; echo "cfg3 graph rc=$?"
timeout 400 python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager --ragged > $O/bench_cfg3_ragged.json 2> $O/bench_cfg3_ragged.err; echo "cfg3 ragged rc=$?"
for t in 16 64 128; do
  CUDA_VISIBLE_DEVICES= timeout 300 python bench.py --impl reference --workload cfg3 --steps 3 --warmup 1 --cpu-threads $t > $O/cpu_threads_$t.json 2> $O/cpu_threads_$t.err
done
VIMA_B200_EPI_PREFETCH=1 GB_REPS=5 timeout 300 python tools/gemm_bench.py > $O/gemm_bench_pf1.txt 2>&1
VIMA_B200_EPI_PREFETCH=0 GB_REPS=5 timeout 300 python tools/gemm_bench.py > $O/gemm_bench_pf0.txt 2>&1
AB_B=256 timeout 200 python tools/attn_bench.py > $O/attn_bench_tc.txt 2>&1
VIMA_B200_ATTN=mma AB_B=256 timeout 200 python tools/attn_bench.py > $O/attn_bench_mma.txt 2>&1
ls -la $O
