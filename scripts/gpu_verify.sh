#!/bin/bash
# one-box verification: the whole -m gpu suite, smoke(), and the driver's bench command line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/verify_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/verify_pytest.log
tail -4 gpurun_out/verify_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/verify_bench_reference.json 2> gpurun_out/verify_bench_reference.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/verify_bench.json')); r=json.load(open('gpurun_out/verify_bench_reference.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'sync', (d['roofline']['sync'] or {}).get('frac'), 'reference arm', r['value'], r['cpu_baseline']['kind'])
PY
