#!/bin/bash
# evidence run with the tuned defaults: tests, bench (all single-GPU workloads), ncu launch list + full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
for w in stress_100k_d8 stress_1m_d16 p2p_1m_maxpred8 stress_10m_d32; do
  timeout 600 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_particles_program -s 16 -c 2 -f -o gpurun_out/prof_final python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_image_tma -c 4 -f -o gpurun_out/prof_tma python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tma.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f cpu=%s ok=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('cpu_baseline',{}).get('value'),d['synctest_consistent']))
PY
