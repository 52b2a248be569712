#!/bin/bash
# Round 2, end-of-round record on one GPU: GPU tests, smoke, both bench arms, launch list, ncu captures of the hot kernels, sanitizers, config-5 sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-config3 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bands_fused -s 1 -c 1 -f -o gpurun_out/prof_extract python scripts/prof_run.py --what extract --clips 10000 --reps 2 > gpurun_out/ncu_extract.log 2>&1; echo "ncu bands rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:haar_select32 -s 1 -c 1 -f -o gpurun_out/prof_select python scripts/prof_run.py --what extract --clips 10000 --reps 2 > gpurun_out/ncu_select.log 2>&1; echo "ncu select rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_fast -s 3 -c 1 -f -o gpurun_out/prof_search python scripts/prof_run.py --what search --db-clips 1000000 > gpurun_out/ncu_search.log 2>&1; echo "ncu search rc=$?"
for t in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $t python scripts/sanitize_run.py"; timeout 900 compute-sanitizer --tool $t python scripts/sanitize_run.py 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20; done > gpurun_out/sanitizer.txt 2>&1; tail -6 gpurun_out/sanitizer.txt
timeout 600 python scripts/config5_sweep.py --out gpurun_out/config5_sweep.json > gpurun_out/config5_sweep.log 2>&1; echo "config5 rc=$?"
timeout 300 python scripts/prof_run.py --what resample --clips 10000 --reps 3 > gpurun_out/resample.log 2>&1; tail -2 gpurun_out/resample.log
ls -la gpurun_out
