#!/bin/bash
# Quick GPU check of a search-kernel change: the search / database parity tests, then the search legs of the bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "search or database or few or regular or merge or group or topk or compare" > gpurun_out/pytest_search.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_search.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-config3 > gpurun_out/bench_search.json 2> gpurun_out/bench_search.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_search.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_search.json"))
s = j["search"]; print("search", s["ms_per_step"], s["kernel_ms"], s["value"], s["topk_sha256"], s.get("single_query_ms"))
print("config5", j.get("config5", {}).get("kernel_ms"))
print("extract", j["ms_per_step"], j["roofline"]["kernel_ms"])
PY
