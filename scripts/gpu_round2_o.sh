#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -4 gpurun_out/r02o_pytest.log
timeout 900 python bench.py --workload p2p_1m_maxpred8 --steps 100 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/r02_bench_p2p_1m_maxpred8.json 2> gpurun_out/r02o_p2p.err; echo "p2p rc=$?"
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02_generic_world.json 2> gpurun_out/r02o_generic.err; echo "generic rc=$?"; cat gpurun_out/r02_generic_world.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02o_bench20.json 2> gpurun_out/r02o_bench20.err; echo "bench20 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o_bench20.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['sync'], d['roofline']['isolated'])
d=json.load(open('gpurun_out/r02_bench_p2p_1m_maxpred8.json'))
print('p2p', d['value'], d['e2e']['value'], d['roofline']['sync'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
