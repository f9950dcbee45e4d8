#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r02_bench_n1_s20.json 2> gpurun_out/r02_bench_n1_s20.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n1.json","gpurun_out/r02_bench_n1_s20.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["value"], d["ms_per_step"], d["stages_ms"], d["clocks"], d["host"], d["e2e"]["ms_per_step"])
PY
