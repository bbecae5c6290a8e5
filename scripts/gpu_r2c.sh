#!/usr/bin/env bash
# round 2, call C: full GPU test-suite (new tests included), PDL / 2-tiles-per-CTA experiments, every bench config
TAG=r02c
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-220
TDS_B200_TPC=2 python -m pytest tests/test_parity_gpu.py -q -m gpu --tb=line -k "laikago or ant or ragged or graph" > gpurun_out/${TAG}_pytest_tpc2.log 2>&1; echo "pytest tpc2 rc=$?"; tail -4 gpurun_out/${TAG}_pytest_tpc2.log | cut -c1-220
TDS_B200_PDL=1 python -m pytest tests/test_parity_gpu.py -q -m gpu --tb=line -k "laikago or ant or ragged or graph or rollout" > gpurun_out/${TAG}_pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"; tail -4 gpurun_out/${TAG}_pytest_pdl.log | cut -c1-220
run() { # name env... (BARGS from the environment)
  name=$1; shift
  env "$@" python bench.py --gpus 1 --steps 500 --warmup 50 --no-cpu-baseline $BARGS > gpurun_out/${TAG}_bench_$name.json 2>> gpurun_out/${TAG}_bench.err
}
BARGS="--envs 4096" run 4096_base TDS_B200_PDL=0
BARGS="--envs 4096" run 4096_pdl TDS_B200_PDL=1
BARGS="--envs 4096 --no-graph" run 4096_pdl_nograph TDS_B200_PDL=1
BARGS="--envs 4096 --no-graph" run 4096_base_nograph TDS_B200_PDL=0
for E in 16384 65536; do
  BARGS="--envs $E" run ${E}_tpc1 TDS_B200_TPC=1
  BARGS="--envs $E" run ${E}_tpc2 TDS_B200_TPC=2
  BARGS="--envs $E" run ${E}_tpc2_pdl TDS_B200_TPC=2 TDS_B200_PDL=1
done
BARGS="--envs 8192" run 8192_tpc1 TDS_B200_TPC=1
BARGS="--envs 8192" run 8192_tpc2 TDS_B200_TPC=2
for C in cartpole64 pendulum5_fd sphere2_16384 humanoid4096 humanoid4096_spring; do
  python bench.py --config $C --steps 100 --warmup 10 > gpurun_out/${TAG}_cfg_$C.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')) + sorted(glob.glob('gpurun_out/${TAG}_cfg*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'frac %.4f' % d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/${TAG}_bench.err
