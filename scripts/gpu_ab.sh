#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json
timeout 200 python -m pytest tests/test_gpu_parity_particles.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 60 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; }
for rep in 1 2; do
run v2_m1_dyn_$rep BGR_TUNE_MINB=1
run v2_m2_dyn_$rep BGR_TUNE_MINB=2
run v2_m8_dyn_$rep BGR_TUNE_MINB=8
done
run v2_m1_static BGR_TUNE_MINB=1 BGR_TUNE_DYNAMIC=0
run v2_m2_static BGR_TUNE_MINB=2 BGR_TUNE_DYNAMIC=0
run v4_m1_dyn BGR_TUNE_VEC=4 BGR_TUNE_MINB=1
run v4_m2_dyn BGR_TUNE_VEC=4 BGR_TUNE_MINB=2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-200:])
PY
