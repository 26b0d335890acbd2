#!/bin/bash
# quick confidence run on a GPU box: parity tests, then one short bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/q_bench.json') if l.startswith('{')][0])
    print("ms=%.4f frac=%.3f e2e=%.0f ok=%s launches=%s" % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'],
                                                        d['synctest_consistent'], d['gpu_launches']))
except Exception as e:
    print("FAILED", e, open('gpurun_out/q_bench.err').read()[-600:])
PY
