#!/bin/bash
# round-2 GPU call 9: SCUNet engine, window attention generalisation, experiment-code removal regression
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engines.py -m gpu -q -s -k "not shapes" > gpurun_out/r02_tests9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests9.log; grep -n "scunet\|rrdbnet\|FAILED\|Error\|passed\|failed" gpurun_out/r02_tests9.log | tail -14
