#!/bin/bash
# multi-GPU bench: N ranks.   usage: tools/gpu_run_n.sh N [cfg4] [collective]
#   default: cfg-2 weak scaling, results exchanged by the refit kernel's NVLink peer stores
#   cfg4:        additionally BASELINE configs[3] (B = 128, 720x540, K = 17) sharded over the N ranks
#   collective:  additionally the one-all_gather-per-step path (--quick) for comparison
N=$1; shift
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_bench_n${N}_peer.json 2> gpurun_out/r02_bench_n${N}_peer.err; echo "peer rc=$?"
for opt in "$@"; do
  if [ "$opt" = "cfg4" ]; then
    timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --workload cfg4 --quick > gpurun_out/r02_bench_n${N}_cfg4_peer.json 2> gpurun_out/r02_bench_n${N}_cfg4_peer.err; echo "cfg4 rc=$?"
  fi
  if [ "$opt" = "collective" ]; then
    timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --gather collective --quick > gpurun_out/r02_bench_n${N}_collective.json 2> gpurun_out/r02_bench_n${N}_collective.err; echo "collective rc=$?"
  fi
done
tail -3 gpurun_out/r02_bench_n${N}_peer.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_n${N}_*.json")):
    try:
        l=[x for x in open(f) if x.startswith("{")][-1]
        d=json.loads(l); print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), d.get("gather_check"), d["config"]["workload"][:12], d["stages_ms"])
    except Exception as e: print(f, "no line:", e)
PY
