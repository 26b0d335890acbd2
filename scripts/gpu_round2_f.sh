#!/bin/bash
mkdir -p gpurun_out
SWEEP_EXTRA='[{"BGR_TUNE_STAGGER_NS": 400}, {"BGR_TUNE_STAGGER_NS": 800}, {"BGR_TUNE_STAGGER_NS": 1500}, {"BGR_TUNE_STAGGER_NS": 3000}, {"BGR_TUNE_STAGGER_NS": 6000}]' timeout 900 python scripts/sync_sweep.py stress_1m_d8 > gpurun_out/r02f_sweep.jsonl 2> gpurun_out/r02f_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02f_sweep.jsonl | cut -c1-420
timeout 900 python bench.py --steps 300 --warmup 5 --trace-out gpurun_out/r02f_timeline.csv > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_python_caller']['value'], d['timeline'])
print(d['roofline']['sync'], d['roofline']['e2e'])
print(d.get('snapshot_save_restore'), d.get('snapshot_save_restore_10m'))
PY
