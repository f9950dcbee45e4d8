#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt
tail -8 gpurun_out/r02_pytest_gpu.txt
timeout 120 tools/pcie_probe.bin > gpurun_out/r02_pcie_probe.txt 2>&1; cat gpurun_out/r02_pcie_probe.txt
timeout 300 python tools/tune_vote.py --variants 0 --chunks 0,1,2 > gpurun_out/r02_gather_modes.txt 2>&1; cat gpurun_out/r02_gather_modes.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/e2e_sweep.py > gpurun_out/r02_e2e_sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pvb -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k regex:pvb -c 60 --csv --log-file gpurun_out/r02_traffic_warm.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_traffic.log 2>&1
timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k regex:pvb -c 24 --csv --log-file gpurun_out/r02_traffic_cold.csv python tools/prof_step.py --steps 3 > gpurun_out/r02_ncu_traffic_cold.log 2>&1
for k in mask_bits gather_kernel vote_kernel refit_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r02_$k python tools/prof_step.py --steps 3 > gpurun_out/r02_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
head -c 1200 gpurun_out/r02_bench_n1.json
