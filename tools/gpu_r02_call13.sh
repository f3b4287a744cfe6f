#!/bin/bash
# round-2 GPU call 13 (8 GPUs): the bench line at N=8 exactly as the driver launches it (sharded tiles, batch units, row-sharded VAE)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02_bench13_n8.json 2> gpurun_out/r02_bench13_n8.err
echo "bench rc=$?"; tail -4 gpurun_out/r02_bench13_n8.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench13_n8.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','phases_ms')}); print('e2e',d['e2e']['value'])
print('b4', d['batch4_512']['value'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['tile_forwards_per_rank'], t['allgather_ms_per_step'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['forwards_per_step_rank0'], v['allgather_ms_per_step'], v['phases_ms_rank0'])
PY
