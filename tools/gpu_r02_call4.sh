#!/bin/bash
# round-2 GPU call 4: changed-kernel tests, slow-test diagnosis, bench, forward A/B + large-batch census, ncu --set full evidence
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engines.py -m gpu -q -s -k "not shapes" > gpurun_out/r02_tests4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests4.log; tail -4 gpurun_out/r02_tests4.log
DBIR_TEST_FAULTDUMP=30 timeout 200 python -m pytest "tests/test_gpu_engines.py::test_cldm_small_vs_oracle_shapes" -m gpu -q -s -k "32-2" > gpurun_out/r02_slowtest.log 2>&1
echo "slowtest rc=$?"; grep -n "rel rms\|File\|passed\|failed" gpurun_out/r02_slowtest.log | head -40
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
echo "bench rc=$?"; head -c 1800 gpurun_out/r02_bench4.json; echo
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census4.csv
AB=gpurun_out/r02_ab4.jsonl; : > $AB
run() { echo "== $*"; env "${@:2}" timeout 400 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab4.err; tail -1 $AB | cut -c1-300; }
EXTRA="" run grouped X=1
run ragged DBIR_GEMM_RAGGED_WIDE=1
run ptmem DBIR_LIB_TAG=ptmem
run earlyb DBIR_LIB_TAG=earlyb
run nopdl DBIR_PDL=0
EXTRA="--census --nb=14" run nb14 X=1
EXTRA="--census --nb=28" run nb28 X=1
# ncu --set full evidence (CSV exported on the box; reports are too big to bring back)
prof() { # name, kernel regex, skip, count
  timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 python tools/profile_image.py 2 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1.csv 2>/dev/null; echo "ncu $1 rc=$? rows=$(wc -l < gpurun_out/r02_ncu_$1.csv)"
}
prof swin "swin_window_attn|layernorm_kernel|swin_stem" 0 6
prof norm "gn_apply|gn_finalize|gn_stats" 0 14
prof normloop "gn_apply|gn_finalize|layernorm_kernel" 220 12
prof small "sampler_step|conv3x3_small|wavelet|linear_f32|im2col|softmax_rows|upsample2x" 0 16
prof vaeconv "gemm_tc_kernel" 215 14
timeout 500 ncu --set full --clock-control none --nvtx --nvtx-include "dbir_fwd/" -k regex:"gemm_tc_kernel|attn_fwd_kernel" -c 60 -f -o /tmp/fwd python tools/profile_forward.py forward 2 > gpurun_out/r02_ncu_fwd.log 2>&1
ncu -i /tmp/fwd.ncu-rep --page raw --csv > gpurun_out/r02_ncu_fwd.csv 2>/dev/null; echo "ncu fwd rows=$(wc -l < gpurun_out/r02_ncu_fwd.csv)"
ls -la /tmp/*.ncu-rep
