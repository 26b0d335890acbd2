#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pf_*.json
timeout 600 python -m pytest tests/test_gpu_parity_particles.py -m gpu -x -q 2>&1 | tail -3
run() { name=$1; n=$2; steps=$3; shift 3; env "$@" timeout 120 python bench.py --entities $n --steps $steps --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/pf_$name.json 2>gpurun_out/pf_$name.err; }
for n in 50000 100000 200000 400000; do
for v in 1 2 4; do run n${n}_v$v $n 2000 BGR_TUNE_VEC=$v; done
done
run n100000_v1_m1 100000 2000 BGR_TUNE_VEC=1 BGR_TUNE_MINB=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/pf_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-300:])
PY
