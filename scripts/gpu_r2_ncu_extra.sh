#!/bin/bash
# ncu captures of the two remaining kernels (few-query search, resampler) with the final code.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_few -s 1 -c 1 -f -o gpurun_out/prof_few python scripts/prof_run.py --what search --db-clips 1000000 --queries 1 > gpurun_out/ncu_few.log 2>&1; echo "ncu few rc=$?"
timeout 800 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 1 -c 1 -f -o gpurun_out/prof_resample python scripts/prof_run.py --what resample --clips 10000 --reps 2 > gpurun_out/ncu_resample.log 2>&1; echo "ncu resample rc=$?"
timeout 300 python scripts/search_latency.py > gpurun_out/search_latency.log 2>&1; tail -6 gpurun_out/search_latency.log
