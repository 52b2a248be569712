#!/bin/bash
# Round 2, 8-GPU record with the final kernels: both bench arms as the driver launches them.
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-600
