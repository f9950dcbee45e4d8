#!/bin/bash
# compute-sanitizer passes over the GPU test-suite with time boxes that finish.  Full logs -> gpurun_out/
# memcheck: all kernels incl. pnp.cu, exchange.cu and the final vote.cu; racecheck: shared-memory hazards of the kernels
# that use shared memory; initcheck: reads of uninitialised device memory (workspace layout, exchange rings).
set -x
mkdir -p gpurun_out
SEL="tests/test_gpu_kernels.py tests/test_gpu_exchange.py tests/test_gpu_pnp.py tests/test_gpu_zz_p3p.py tests/test_gpu_decode.py tests/test_gpu_golden.py"
LAYER='tests/test_gpu_layer.py -k "not full_size and not graph"'
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 0 --error-exitcode 99 python -m pytest $SEL -x -q -m gpu > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
timeout 600 bash -c "compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest $LAYER -x -q -m gpu" > gpurun_out/r02_sanitizer_memcheck_layer.txt 2>&1; echo "memcheck layer rc=$?" >> gpurun_out/r02_sanitizer_memcheck_layer.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_exchange.py tests/test_gpu_pnp.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 99 python -m pytest tests/test_gpu_golden.py tests/test_gpu_exchange.py tests/test_gpu_pnp.py tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_sanitizer_initcheck.txt 2>&1; echo "initcheck rc=$?" >> gpurun_out/r02_sanitizer_initcheck.txt
tail -4 gpurun_out/r02_sanitizer_*.txt
