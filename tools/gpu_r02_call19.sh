#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_engines.py -m gpu -q -s -k "cldm_full_config or rrdbnet_full or scunet_full or vae_full or (swinir_full and 256)" > gpurun_out/r02_tests19.log 2>&1
echo "pytest rc=$?"; grep -n "fixture\|passed\|failed\|Error\|error" gpurun_out/r02_tests19.log | tail -10
