#!/bin/bash
# round-2 final single-GPU session: tests, both bench arms, ncu evidence (launch list, DRAM traffic warm/cold, --set full of
# the five kernels of a step), compute-sanitizer passes with time boxes.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.txt; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
K='regex:mask_bits_kernel|thin_gather_kernel|generate_kernel|vote_kernel|refit_kernel|exchange_wait_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 36 --csv --log-file gpurun_out/r02_launches.csv python bench.py --quick --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k "$K" -s 24 -c 18 --csv --log-file gpurun_out/r02_traffic_warm.csv python tools/prof_step.py --steps 8 > gpurun_out/r02_ncu_traffic.log 2>&1
timeout 300 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k "$K" -s 12 -c 12 --csv --log-file gpurun_out/r02_traffic_cold.csv python tools/prof_step.py --steps 5 > gpurun_out/r02_ncu_traffic_cold.log 2>&1
for k in mask_bits_kernel thin_gather_kernel vote_kernel refit_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$k python tools/prof_step.py --steps 4 > gpurun_out/r02_ncu_$k.log 2>&1
done
# ---- compute-sanitizer (full logs kept)
SEL="tests/test_gpu_kernels.py tests/test_gpu_exchange.py tests/test_gpu_pnp.py tests/test_gpu_zz_p3p.py tests/test_gpu_decode.py tests/test_gpu_golden.py"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest $SEL -x -q -m gpu > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_layer.py -x -q -m gpu -k "not full_size and not graph and not large" > gpurun_out/r02_sanitizer_memcheck_layer.txt 2>&1; echo "memcheck layer rc=$?" >> gpurun_out/r02_sanitizer_memcheck_layer.txt
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_exchange.py tests/test_gpu_golden.py tests/test_gpu_pnp.py -x -q -m gpu > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
timeout 420 compute-sanitizer --tool initcheck --error-exitcode 99 python -m pytest tests/test_gpu_golden.py tests/test_gpu_exchange.py tests/test_gpu_pnp.py tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_sanitizer_initcheck.txt 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02_sanitizer_initcheck.txt
tail -3 gpurun_out/r02_sanitizer_*.txt
ls -la gpurun_out/*.ncu-rep
