#!/bin/bash
# ncu evidence pass (single GPU): launch list of the bench command, DRAM traffic of one step with warm and cold caches,
# `--set full` captures of the select-stage kernels, the vote kernel and the refit kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 120 tools/pcie_probe.bin > gpurun_out/r02_pcie_probe.txt 2>&1; tail -9 gpurun_out/r02_pcie_probe.txt
timeout 300 python tools/tune_vote.py --variants 0 --chunks 0,2 > gpurun_out/r02_gather_modes2.txt 2>&1; cat gpurun_out/r02_gather_modes2.txt
K='regex:mask_bits_kernel|thin_gather_kernel|generate_kernel|vote_kernel|refit_kernel|exchange_wait_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 72 --csv --log-file gpurun_out/r02_launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k "$K" -s 24 -c 18 --csv --log-file gpurun_out/r02_traffic_warm.csv python tools/prof_step.py --steps 8 > gpurun_out/r02_ncu_traffic.log 2>&1
timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k "$K" -s 12 -c 12 --csv --log-file gpurun_out/r02_traffic_cold.csv python tools/prof_step.py --steps 5 > gpurun_out/r02_ncu_traffic_cold.log 2>&1
for k in mask_bits_kernel thin_scan_kernel gather_kernel vote_kernel refit_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$k python tools/prof_step.py --steps 4 > gpurun_out/r02_ncu_$k.log 2>&1
  tail -2 gpurun_out/r02_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
