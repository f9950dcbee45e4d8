#!/bin/bash
# round-2 GPU session 1: tests, bench, e2e sweep, launch list, realistic DRAM traffic of the step, L2-granularity experiment
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
tail -5 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/e2e_sweep.py > gpurun_out/r02_e2e_sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -c 120 --csv --log-file gpurun_out/r02_traffic_warm.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_traffic.log 2>&1
for g in 0 32 128; do
  timeout 300 python tools/l2gran.py $g >> gpurun_out/r02_l2gran.txt 2>&1
  timeout 300 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gather_kernel -c 6 --csv --log-file gpurun_out/r02_l2gran_$g.csv python tools/l2gran.py $g > /dev/null 2>&1
done
tail -3 gpurun_out/r02_l2gran.txt
head -c 1500 gpurun_out/r02_bench_n1.json
