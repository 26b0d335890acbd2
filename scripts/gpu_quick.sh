#!/bin/bash
# parity + a few bench variants (prints one line each)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/q_$name.json 2>gpurun_out/q_$name.err; }
run default A=1
run default2 A=1
run v4m1 BGR_TUNE_VEC=4 BGR_TUNE_MINB=1
run v2m1 BGR_TUNE_MINB=1
run static BGR_TUNE_DYNAMIC=0
run bps3 BGR_TUNE_BPS=3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/q_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-300:])
PY
