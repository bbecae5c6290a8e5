#!/usr/bin/env bash
TAG=r02d
mkdir -p gpurun_out
python scripts/dbg_ars.py 2>&1 | tail -8
python -m pytest tests/test_parity_gpu.py -q -m gpu --tb=line -k "full_size or v2_abi or ars_iteration" 2>&1 | tail -5 | cut -c1-200
for C in humanoid4096 humanoid4096_spring; do
  python bench.py --config $C --steps 100 --warmup 10 > gpurun_out/${TAG}_cfg_$C.json 2>> gpurun_out/${TAG}_bench.err
done
run() { name=$1; shift; env "$@" python bench.py --gpus 1 --steps 500 --warmup 50 --no-cpu-baseline $BARGS > gpurun_out/${TAG}_bench_$name.json 2>> gpurun_out/${TAG}_bench.err; }
BARGS="--envs 4096" run 4096_base A=1
BARGS="--envs 4096" run 4096_rcp TDS_B200_LIB=$PWD/tiny-differentiable-simulator_b200/libtds_b200_rcp.so
BARGS="--envs 65536" run 65536_base A=1
BARGS="--envs 65536" run 65536_rcp TDS_B200_LIB=$PWD/tiny-differentiable-simulator_b200/libtds_b200_rcp.so
TDS_B200_LIB=$PWD/tiny-differentiable-simulator_b200/libtds_b200_rcp.so python -m pytest tests/test_parity_gpu.py -q -m gpu --tb=line -k "laikago or ant or golden or ragged" 2>&1 | tail -4 | cut -c1-200
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'frac %.4f' % d['roofline']['frac'], d['roofline'].get('kernel','')[:30])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/${TAG}_bench.err
