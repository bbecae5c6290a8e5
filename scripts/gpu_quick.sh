#!/usr/bin/env bash
# Short gpurun call while iterating on a kernel: GPU tests, phase clocks, bench of the default kernel and of the
# kernels named in $2.. for comparison (no CPU baseline), optional racecheck of a small run.
# Usage: scripts/gpu_quick.sh <tag> [kernel ...]
TAG="${1:-q}"; shift
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases.txt
python bench.py --gpus 1 --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
for K in "$@"; do
  TDS_B200_KERNEL=$K python bench.py --gpus 1 --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/${TAG}_bench_$K.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
if [ -n "$RACECHECK" ]; then
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/race_small.py > gpurun_out/${TAG}_racecheck.log 2>&1
  tail -15 gpurun_out/${TAG}_racecheck.log
fi
if [ -n "$NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:tds_step -s 262 -c 1 -f -o gpurun_out/${TAG}_step_full \
      python scripts/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/${TAG}_step_full.ncu-rep
fi
