#!/usr/bin/env bash
# Multi-GPU evidence (one node, N GPUs visible): weak scaling, the {reward, done} exchange inside the step graph (same
# stream / overlapped), strong scaling.   Usage: scripts/gpu_multi.sh <tag> "<N list>"      e.g.  r02 "2 4 8"
TAG="${1:-r02}"; NS="${2:-2}"
mkdir -p gpurun_out
run() { # N name args...
  N=$1; name=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $N --steps 500 --warmup 50 --no-cpu-baseline "$@" > gpurun_out/${TAG}_N${N}_${name}.json 2>> gpurun_out/${TAG}_multi.err
  echo "N=$N $name rc=$?"
}
for N in $NS; do
  run $N weak
  run $N gather --gather-reward
  run $N gather_overlap --gather-reward --gather-overlap
  run $N strong --strong
  run $N strong_gather_overlap --strong --gather-reward --gather-overlap
done
NMAX=$(echo $NS | awk '{print $NF}')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29417 \
    bench.py --gpus $NMAX --config humanoid4096 --steps 100 --warmup 10 > gpurun_out/${TAG}_N${NMAX}_humanoid.json 2>> gpurun_out/${TAG}_multi.err
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_N*_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['scaling'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/${TAG}_multi.err
