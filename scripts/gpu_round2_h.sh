#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -6 gpurun_out/r02h_pytest.log
SWEEP_EXTRA='[{"BGR_TUNE_SUB": 512}, {"BGR_TUNE_SUB": 128}]' timeout 900 python scripts/sync_sweep.py stress_100k_d8 > gpurun_out/r02h_sweep.jsonl 2> gpurun_out/r02h_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02h_sweep.jsonl | cut -c1-520
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02h_generic_world.json 2> gpurun_out/r02h_generic_world.err; echo "generic rc=$?"; cat gpurun_out/r02h_generic_world.json; tail -3 gpurun_out/r02h_generic_world.err
