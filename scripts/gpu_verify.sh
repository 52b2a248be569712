#!/bin/bash
# Short one-GPU check: parity tests, smoke, both bench arms (a few GPU-minutes).  Usage (under gpurun): bash scripts/gpu_verify.sh [full]
# "full" adds the config-3 run at full size, the resampler timing + ncu capture and the compute-sanitizer passes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --microbench > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
if [ "$1" = "full" ]; then
  timeout 600 python bench.py --clips 120000 --no-e2e --no-search --no-cpu --steps 3 --warmup 3 > gpurun_out/bench_config3_n1.json 2> gpurun_out/bench_config3_n1.err; echo "config 3 rc=$?"
  timeout 300 python scripts/prof_run.py --what resample --clips 10000 --reps 5 > gpurun_out/resample.log 2>&1; tail -2 gpurun_out/resample.log
  timeout 800 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 1 -c 1 -f -o gpurun_out/prof_resample python scripts/prof_run.py --what resample --clips 10000 --reps 2 > gpurun_out/ncu_resample.log 2>&1; echo "ncu resample rc=$?"
  for t in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $t python scripts/sanitize_run.py"; timeout 900 compute-sanitizer --tool $t python scripts/sanitize_run.py 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20; done > gpurun_out/sanitizer.txt 2>&1; tail -6 gpurun_out/sanitizer.txt
fi
