#!/usr/bin/env bash
# Round-2 evidence in one gpurun call (1 GPU): GPU tests, smoke, phase clocks, the headline bench + reference arm, every
# BASELINE config, ncu launch list WITHOUT the per-launch cache flush (kernel durations comparable with the CUDA-event
# figure), one full ncu capture of the step kernel, racecheck of a small run.   Usage: scripts/gpu_round2.sh <tag>
TAG="${1:-r02}"
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; echo nproc $(nproc); python -c "import os;print('affinity',len(os.sched_getaffinity(0)))"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg; } > gpurun_out/${TAG}_box.txt 2>&1
python -m pytest tests/ -q -m gpu --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/${TAG}_smoke.log | tail -2
python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases.txt | head -14
python bench.py --gpus 1 --steps 1000 --warmup 100 > gpurun_out/${TAG}_bench_N1.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_N1_driver_like.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
for E in 16384 65536; do
  python bench.py --envs $E --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${TAG}_bench_${E}.json 2>> gpurun_out/${TAG}_bench.err
done
for C in cartpole64 pendulum5_fd sphere2_16384 humanoid4096 humanoid4096_spring; do
  python bench.py --config $C --steps 100 --warmup 10 > gpurun_out/${TAG}_cfg_$C.json 2>> gpurun_out/${TAG}_bench.err
  [ "$C" != humanoid4096_spring ] && python bench.py --impl reference --config $C --steps 20 --warmup 5 > gpurun_out/${TAG}_cfgref_$C.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')) + sorted(glob.glob('gpurun_out/${TAG}_cfg*.json')):
    try:
        d = json.load(open(f)); r = d.get('roofline') or {}
        print(f.split('/')[-1], 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'frac', r.get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
grep -c tds_step gpurun_out/${TAG}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:tds_step -s 262 -c 1 -f -o gpurun_out/${TAG}_step_full \
    python scripts/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_step_full.ncu-rep
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/race_small.py > gpurun_out/${TAG}_racecheck.log 2>&1
tail -3 gpurun_out/${TAG}_racecheck.log
python scripts/bench_rollout.py > gpurun_out/${TAG}_rollout.txt 2>&1; tail -2 gpurun_out/${TAG}_rollout.txt
