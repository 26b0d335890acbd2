#!/bin/bash
mkdir -p gpurun_out
for e in 125000 250000 500000 1000000 2000000 4000000 8000000; do
  timeout 300 python bench.py --entities $e --steps 300 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/fit_$e.json 2>/dev/null
done
python - <<'PY'
import json
for e in [125000,250000,500000,1000000,2000000,4000000,8000000]:
    d=json.loads([l for l in open(f'gpurun_out/fit_{e}.json') if l.startswith('{')][0])
    print(e, "ms=%.4f"%d['ms_per_step'], "frac=%.3f"%d['roofline']['frac'], "e2e_ms=%.4f"%(d['config']['advances_per_step']/d['e2e']['value']*1e3))
PY
