#!/usr/bin/env bash
TAG=r02h
mkdir -p gpurun_out
V=$PWD/tiny-differentiable-simulator_b200/libtds_b200_dsplit.so
TDS_B200_LIB=$V python -m pytest tests/ -q -m gpu --tb=line 2>&1 | tail -6 | cut -c1-200
run() { name=$1; shift; env "$@" python bench.py --gpus 1 --steps 500 --warmup 50 --no-cpu-baseline $BARGS > gpurun_out/${TAG}_bench_$name.json 2>> gpurun_out/${TAG}_bench.err; }
BARGS="--envs 4096" run 4096_base A=1
BARGS="--envs 4096" run 4096_dsplit TDS_B200_LIB=$V
BARGS="--envs 65536" run 65536_base A=1
BARGS="--envs 65536" run 65536_dsplit TDS_B200_LIB=$V
BARGS="--envs 4096 --steps 20 --warmup 5" run 4096_driverlike A=1
TDS_B200_LIB=$V python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases_dsplit.txt | head -14
TDS_B200_LIB=$V python scripts/bench_ant.py 2>&1 | tail -3
python scripts/bench_ant.py 2>&1 | tail -3
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
