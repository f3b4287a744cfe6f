#!/bin/bash
# round-2 GPU call 16: ncu --set full evidence of the final kernels (normalisation passes, small kernels, forward GEMMs / attention)
mkdir -p gpurun_out
prof() { # name, kernel regex, skip, count
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 python tools/profile_image.py 2 > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1.csv 2>/dev/null; echo "ncu $1 rc=$? rows=$(wc -l < gpurun_out/r02_ncu_$1.csv)"
}
prof norm2 "gn_apply|gn_finalize|gn_stats" 0 16
prof small2 "sampler_step|conv3x3_small|wavelet|linear_f32|im2col|softmax_rows|axpby" 0 16
timeout 400 ncu --set full --clock-control none --nvtx --nvtx-include "dbir_fwd/" -k regex:"gemm_tc_kernel|attn_fwd_kernel" -c 70 -f -o /tmp/fwd2 python tools/profile_forward.py forward 2 > gpurun_out/r02_ncu_fwd2.log 2>&1
ncu -i /tmp/fwd2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_fwd2.csv 2>/dev/null; echo "ncu fwd rows=$(wc -l < gpurun_out/r02_ncu_fwd2.csv)"
