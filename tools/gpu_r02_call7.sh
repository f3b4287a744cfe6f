#!/bin/bash
# round-2 GPU call 7 (2 GPUs): sharded paths under torchrun -- bit-identity tool (tiles + batch units) and the bench line at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tools/run_tiled_multi.py > gpurun_out/r02_multi7.log 2>&1
echo "multi rc=$?"; grep -n "bit-equal\|Error\|error" gpurun_out/r02_multi7.log | head
DBIR_FULL=1 DBIR_STEPS=3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 tools/run_tiled_multi.py > gpurun_out/r02_multi7_full.log 2>&1
echo "multi full rc=$?"; grep -n "bit-equal\|Error\|error" gpurun_out/r02_multi7_full.log | head
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench7_n2.json 2> gpurun_out/r02_bench7_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench7_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench7_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','phases_ms')}); print('e2e',d['e2e']['value'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['tiles_per_rank'], t['allgather_ms_per_step'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['forwards_per_step_rank0'], v['allgather_ms_per_step'], v['phases_ms_rank0'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29544 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_bench7_ref_n2.json 2> gpurun_out/r02_bench7_ref_n2.err
echo "ref rc=$?"; head -c 600 gpurun_out/r02_bench7_ref_n2.json
