#!/usr/bin/env bash
# round 2, call A: baseline of the shipped kernel on this round's box + instruction-cache probe + host core facts
TAG=r02a
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; echo nproc $(nproc); python -c "import os;print('affinity',len(os.sched_getaffinity(0)),'cpu_count',os.cpu_count())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg; lscpu | head -20; } > gpurun_out/${TAG}_box.txt 2>&1
./scripts/icache_probe.bin > gpurun_out/${TAG}_icache_probe.txt 2>&1; tail -50 gpurun_out/${TAG}_icache_probe.txt
python scripts/phase_profile.py 2>&1 | tee gpurun_out/${TAG}_phases.txt
for E in 4096 16384 65536; do
  python bench.py --gpus 1 --steps 500 --warmup 50 --envs $E --no-cpu-baseline > gpurun_out/${TAG}_bench_$E.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d = json.load(open(f)); print(f, 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
python -m pytest tests/ -q -m gpu --tb=short -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
cat gpurun_out/${TAG}_box.txt | head -8
