#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_zz_inference_cli.py -m gpu -q -s -k "custom" > gpurun_out/r02_tests18.log 2>&1
echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error\|error" gpurun_out/r02_tests18.log | tail -8
