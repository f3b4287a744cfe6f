#!/bin/bash
# round-2 GPU call 2: grouped-twin correctness + forward A/B experiments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engines.py tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02_tests2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests2.log
tail -4 gpurun_out/r02_tests2.log
AB=gpurun_out/r02_ab.jsonl
: > $AB
run() { echo "== $*"; env "${@:2}" timeout 600 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab.err; tail -1 $AB | cut -c1-400; }
EXTRA="--census" run grouped X=1
run ungrouped DBIR_GROUP_TWINS=0
run grouped_wide0 DBIR_GEMM_WIDE=0
run ungrouped_wide0 DBIR_GROUP_TWINS=0 DBIR_GEMM_WIDE=0
run grouped_earlyb DBIR_LIB_TAG=earlyb
run grouped_ptmem DBIR_LIB_TAG=ptmem
run grouped_ragged DBIR_GEMM_RAGGED_WIDE=1
run grouped_nopdl DBIR_PDL=0
EXTRA="--census --nb=28" run grouped_nb28 X=1
