#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -8 gpurun_out/r02d_pytest.log
timeout 900 python - > gpurun_out/r02d_snap.json 2> gpurun_out/r02d_snap.err <<'PY'
import json, bench
out = {"1m": bench.snapshot_bench(1_000_000, 9, 0), "10m": bench.snapshot_bench(10_000_000, 9, 0, iters=20)}
print(json.dumps(out))
PY
echo "snap rc=$?"; cat gpurun_out/r02d_snap.json; tail -3 gpurun_out/r02d_snap.err
SWEEP_EXTRA='[]' timeout 600 python scripts/sync_sweep.py > gpurun_out/r02d_sweep.jsonl 2> gpurun_out/r02d_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02d_sweep.jsonl
