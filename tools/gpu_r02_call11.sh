#!/bin/bash
# round-2 GPU call 11: GEGLU epilogue A/B (packed / smem-bias vs the previous scalar / shuffle build), CLI cases
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "geglu or gemm" > gpurun_out/r02_tests11.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests11.log; tail -3 gpurun_out/r02_tests11.log
timeout 600 python -m pytest tests/test_gpu_engines.py -m gpu -q -s -k "cldm_full or golden" > gpurun_out/r02_tests11b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests11b.log; grep -n "rel rms\|passed\|failed" gpurun_out/r02_tests11b.log | tail -8
AB=gpurun_out/r02_ab11.jsonl; : > $AB
run() { echo "== $*"; env "${@:2}" timeout 500 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab11.err; tail -1 $AB | cut -c1-420; }
EXTRA="--census" run packed X=1
EXTRA="--census" run scalar DBIR_LIB_TAG=geglusc
EXTRA="--census --nb=28" run det28_packed DBIR_DETERMINISTIC=1
EXTRA="--census --nb=28" run det28_scalar DBIR_DETERMINISTIC=1 DBIR_LIB_TAG=geglusc
timeout 600 python -m pytest tests/test_gpu_zz_inference_cli.py -m gpu -q -s > gpurun_out/r02_tests11c.log 2>&1
echo "cli rc=$?"; tail -3 gpurun_out/r02_tests11c.log
