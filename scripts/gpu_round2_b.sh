#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -8 gpurun_out/r02b_pytest.log
SWEEP_EXTRA='[{"BGR_TUNE_POLL": 0}]' timeout 600 python scripts/sync_sweep.py > gpurun_out/r02b_sweep.jsonl 2> gpurun_out/r02b_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02b_sweep.jsonl
