#!/bin/bash
# round-2 evidence run (1 GPU): bench of every single-GPU workload, device timeline, ncu launch list + full captures.
# Raw reports stay in gpurun_out/ (scratch); tools/ncu_summary.py distils them into profiles/ afterwards (on the CPU box).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 300 --warmup 5 --trace-out gpurun_out/r02_timeline.csv > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
for w in stress_100k_d8 stress_1m_d16 p2p_1m_maxpred8 stress_10m_d32; do
  timeout 900 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; echo "$w rc=$?"
done
# launch list of a bench run (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 120 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-snapshot-bench > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
# full captures: fused kernel (headline workload), TMA copy at 10M (out of L2), generic program (presence world)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_particles_program -s 16 -c 2 -f -o gpurun_out/r02_prof_fused python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-snapshot-bench > gpurun_out/r02_ncu_fused.log 2>&1; echo "fused rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_particles_program -s 16 -c 2 -f -o gpurun_out/r02_prof_fused_100k python bench.py --workload stress_100k_d8 --steps 5 --warmup 3 --no-cpu-baseline --no-snapshot-bench > gpurun_out/r02_ncu_fused_100k.log 2>&1; echo "fused 100k rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_image_tma -s 6 -c 4 -f -o gpurun_out/r02_prof_tma python -c "import bench; bench.snapshot_bench(10_000_000, 9, 0, iters=3)" > gpurun_out/r02_ncu_tma.log 2>&1; echo "tma rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_program -s 20 -c 2 -f -o gpurun_out/r02_prof_generic python scripts/generic_world_bench.py 100000 24 > gpurun_out/r02_ncu_generic.log 2>&1; echo "generic rc=$?"
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02_generic_world.json 2> gpurun_out/r02_generic_world.err; echo "generic bench rc=$?"; cat gpurun_out/r02_generic_world.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f sync=%s ok=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],(d['roofline'].get('sync') or {}).get('frac'),d['synctest_consistent']))
PY
