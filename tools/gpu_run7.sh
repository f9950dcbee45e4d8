#!/bin/bash
# round-2: the other BASELINE configurations with the final kernels, and the production (planar NCHW view) layout
set -x
mkdir -p gpurun_out
timeout 900 python tools/sweep.py --B 64 --steps 10 > gpurun_out/r02_sweep.md 2> gpurun_out/r02_sweep.err; tail -5 gpurun_out/r02_sweep.md
timeout 300 python bench.py --layout planar --steps 100 --warmup 5 > gpurun_out/r02_bench_n1_planar.json 2> gpurun_out/r02_bench_n1_planar.err
timeout 300 python bench.py --impl reference --layout planar --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_planar.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n1_planar.json","gpurun_out/r02_bench_reference_planar.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["value"], d["ms_per_step"], d.get("stages_ms"), (d.get("extras") or {}).get("three_streams"))
PY
