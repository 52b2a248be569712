#!/bin/bash
# Round 2, first GPU call: full-size parity record, GPU tests with figures, baseline bench of the round-1 kernels.
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python tests/parity_record.py --out gpurun_out/r02_parity.json > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/parity.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
