#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -4 gpurun_out/r02i_pytest.log
SWEEP_EXTRA='[]' timeout 900 python scripts/sync_sweep.py > gpurun_out/r02i_sweep.jsonl 2> gpurun_out/r02i_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02i_sweep.jsonl | cut -c1-520
