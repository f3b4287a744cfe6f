#!/bin/bash
# round-2 GPU call 3: full -m gpu suite, the N=1 bench line, launch list of one steady-state image
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r02_tests3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests3.log
tail -5 gpurun_out/r02_tests3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
echo "bench rc=$?"
tail -3 gpurun_out/r02_bench3.err
head -c 3000 gpurun_out/r02_bench3.json
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census3.csv
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/r02_image_launches.csv python tools/profile_image.py 10 > gpurun_out/r02_profile_image.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_profile_image.log; wc -l gpurun_out/r02_image_launches.csv
