#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pf_*.json
for bt in 2 4; do
BENCH_BATCH_TICKS=$bt timeout 120 python bench.py --steps 400 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/pf_bt$bt.json 2>gpurun_out/pf_bt$bt.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/pf_bt$bt.json') if l.startswith('{')][0])
    print("bt$bt ms=%.4f"%d['ms_per_step'], d.get('batched_ticks_experiment'), d['synctest_consistent'])
except Exception as e:
    print("FAILED", e, open('gpurun_out/pf_bt$bt.err').read()[-500:])
PY
done
