#!/bin/bash
# round-2 GPU call 8 (2 GPUs): unit-sharded tiles + row-sharded VAE under torchrun, bench at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tools/run_tiled_multi.py > gpurun_out/r02_multi8.log 2>&1
echo "multi rc=$?"; grep -n "bit-equal\|vae \|Error\|error" gpurun_out/r02_multi8.log | head -20
DBIR_FULL=1 DBIR_STEPS=4 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 tools/run_tiled_multi.py > gpurun_out/r02_multi8_full.log 2>&1
echo "multi full rc=$?"; grep -n "bit-equal\|vae \|Error\|error" gpurun_out/r02_multi8_full.log | head -20
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench8_n2.json 2> gpurun_out/r02_bench8_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench8_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench8_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','phases_ms')}); print('e2e',d['e2e']['value'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['tile_forwards_per_rank'], t['allgather_ms_per_step'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['forwards_per_step_rank0'], v['allgather_ms_per_step'], v['phases_ms_rank0'])
PY
