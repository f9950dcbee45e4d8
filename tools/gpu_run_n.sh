#!/bin/bash
# multi-GPU bench: N ranks, fused peer exchange vs one collective per step.   usage: tools/gpu_run_n.sh N [extra bench args]
N=$1; shift
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/r02_bench_n${N}_peer.json 2> gpurun_out/r02_bench_n${N}_peer.err; echo "peer rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --gather collective --quick "$@" > gpurun_out/r02_bench_n${N}_collective.json 2> gpurun_out/r02_bench_n${N}_collective.err; echo "collective rc=$?"
tail -3 gpurun_out/r02_bench_n${N}_peer.err
python - <<PY
import json
for m in ("peer","collective"):
    try:
        l=[x for x in open(f"gpurun_out/r02_bench_n${N}_{m}.json") if x.startswith("{")][-1]
        d=json.loads(l); print(m, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), d.get("gather_check"), d["config"]["parallelism"][:60], d["stages_ms"])
    except Exception as e: print(m, "no line:", e)
PY
