#!/usr/bin/env bash
# One gpurun call that produces the round's evidence: GPU tests, smoke, phase clocks, bench (graph + no-graph),
# reference arm, ncu launch list of the bench command, one full ncu capture of the step kernel, racecheck.
# Usage: scripts/gpu_round.sh <tag>
TAG="${1:-r01}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/${TAG}_smoke.log | tail -2
python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases.txt
python bench.py --gpus 1 --steps 1000 --warmup 100 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
python bench.py --gpus 1 --steps 1000 --warmup 100 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_bench_nograph.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ['gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_bench_nograph.json','gpurun_out/${TAG}_bench_reference.json']:
    try:
        d=json.load(open(f)); print(f, 'value %.3e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c tds_step gpurun_out/${TAG}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:tds_step -s 262 -c 1 -f -o gpurun_out/${TAG}_step_full \
    python scripts/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_step_full.ncu-rep
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/race_small.py > gpurun_out/${TAG}_racecheck.log 2>&1
tail -3 gpurun_out/${TAG}_racecheck.log
./scripts/ifetch_probe.bin > gpurun_out/${TAG}_ifetch_probe.txt 2>&1; tail -4 gpurun_out/${TAG}_ifetch_probe.txt
python scripts/bench_rollout.py > gpurun_out/${TAG}_rollout.txt 2>&1; tail -2 gpurun_out/${TAG}_rollout.txt
