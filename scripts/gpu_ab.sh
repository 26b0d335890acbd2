#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab_*.json
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 60 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; }
for rep in 1 2; do
run pdl1_$rep A=1
run pdl0_$rep BGR_TUNE_PDL=0
done
run pdl1_100k BGR_X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-200:])
PY
for pdl in 1 0; do BGR_TUNE_PDL=$pdl timeout 60 python bench.py --workload stress_100k_d8 --steps 2000 --warmup 5 --no-cpu-baseline --no-snapshot-bench 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('100k pdl=$pdl ms=%.4f e2e=%.0f'%(d['ms_per_step'],d['e2e']['value']))"; done
