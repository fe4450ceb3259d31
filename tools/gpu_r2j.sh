#!/bin/bash
# round-2 GPU pass J (2 GPUs): the driver's N=2 command: graph replay + NCCL all-gather, gather check, per-rank times
O=gpurun_out/r2j; mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; cut -c1-500 $O/bench_n2.json; tail -5 $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > $O/bench_n2_reference.json 2> $O/bench_n2_reference.err; echo "n2 ref rc=$?"; cut -c1-500 $O/bench_n2_reference.json
timeout 300 python -m pytest tests/test_multi_device_gpu.py -m gpu -q > $O/pytest_multi_device.log 2>&1; tail -3 $O/pytest_multi_device.log
