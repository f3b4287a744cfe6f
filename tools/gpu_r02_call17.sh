#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s -k "reference_whole_run" > gpurun_out/r02_tests17.log 2>&1
echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error" gpurun_out/r02_tests17.log | tail -5
