#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/q_bench.json') if l.startswith('{')][0])
    print("ms=%.4f frac=%.3f e2e=%.0f ok=%s launches=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent'],d['gpu_launches']))
    print("mirror", d.get('e2e_host_mirror'))
except Exception as e:
    print("FAILED", e, open('gpurun_out/q_bench.err').read()[-600:])
PY
timeout 100 python bench.py --workload stress_100k_d8 --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('100k ms=%.4f e2e=%.0f'%(d['ms_per_step'],d['e2e']['value']), d.get('e2e_host_mirror'))"
