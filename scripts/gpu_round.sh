#!/usr/bin/env bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list (+ optional full capture of the step kernel).
# Usage: scripts/gpu_round.sh <tag> [full]
TAG="${1:-r01}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/${TAG}_smoke.log | tail -2
python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases.txt
python bench.py --gpus 1 --steps 1000 --warmup 100 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | head -c 3000; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --gpus 1 --steps 1000 --warmup 100 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_bench_nograph.json 2>> gpurun_out/${TAG}_bench.err
python -c "
import json
for f in ['gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_nograph.json']:
    try:
        d=json.load(open(f)); print(f, 'value %.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c tds_step gpurun_out/${TAG}_launches.csv
if [ "$2" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:tds_step_kernel -s 12 -c 2 -f -o gpurun_out/${TAG}_step_full \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/${TAG}_step_full.ncu-rep
fi
