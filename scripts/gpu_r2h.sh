#!/usr/bin/env bash
# final 1-GPU call of the round: the whole GPU suite on the shipped library, the same on the split-trunk-solve variant, A/B
TAG=r02h
mkdir -p gpurun_out
V=$PWD/tiny-differentiable-simulator_b200/libtds_b200_dsplit.so
python -m pytest tests/ -q -m gpu --tb=line 2>&1 | tail -4 | cut -c1-200 | tee gpurun_out/${TAG}_pytest_default.txt
run() { name=$1; shift; env "$@" python bench.py --gpus 1 --no-cpu-baseline --min-seconds 0.4 $BARGS > gpurun_out/${TAG}_bench_$name.json 2>> gpurun_out/${TAG}_bench.err; }
BARGS="--envs 4096 --steps 20 --warmup 5" run 4096_driverlike A=1
BARGS="--envs 4096 --steps 500 --warmup 50" run 4096_base A=1
BARGS="--envs 4096 --steps 500 --warmup 50" run 4096_dsplit TDS_B200_LIB=$V
BARGS="--envs 65536 --steps 200 --warmup 20" run 65536_base A=1
BARGS="--envs 65536 --steps 200 --warmup 20" run 65536_dsplit TDS_B200_LIB=$V
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
TDS_B200_LIB=$V python -m pytest tests/ -q -m gpu --tb=line 2>&1 | tail -4 | cut -c1-200 | tee gpurun_out/${TAG}_pytest_dsplit.txt
TDS_B200_LIB=$V python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases_dsplit.txt | head -14
tail -3 gpurun_out/${TAG}_bench.err
