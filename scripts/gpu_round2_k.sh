#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -4 gpurun_out/r02k_pytest.log
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02k_generic_world.json 2> gpurun_out/r02k_generic_world.err; echo "generic rc=$?"; cat gpurun_out/r02k_generic_world.json
timeout 600 python scripts/generic_world_bench.py 1000000 60 > gpurun_out/r02k_generic_world_1m.json 2>> gpurun_out/r02k_generic_world.err; echo "generic 1m rc=$?"; cat gpurun_out/r02k_generic_world_1m.json
