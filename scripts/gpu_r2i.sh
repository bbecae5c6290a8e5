#!/usr/bin/env bash
# 2 GPUs, short: the {reward, done} exchange overlapped with the next step (and the same-stream form for the same box)
TAG=r02i
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus 2 --steps 300 --warmup 20 --no-cpu-baseline --min-seconds 0.3 "$@" > gpurun_out/${TAG}_N2_${name}.json 2>> gpurun_out/${TAG}_multi.err
  echo "$name rc=$?"; }
run gather_overlap --gather-reward --gather-overlap
run weak
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_N2_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -n "Error\|error" gpurun_out/${TAG}_multi.err | head -5
