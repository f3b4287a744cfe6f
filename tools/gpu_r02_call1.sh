#!/bin/bash
# round-2 GPU call 1: full -m gpu suite + the N=1 bench line (with the tiled-2048 block)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r02_tests1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests1.log
tail -5 gpurun_out/r02_tests1.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
echo "bench rc=$?"
tail -3 gpurun_out/r02_bench1.err
head -c 1500 gpurun_out/r02_bench1.json
