#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
./tools/hbm_mix_bench > gpurun_out/hbm_mix.json 2>&1; cat gpurun_out/hbm_mix.json
./tools/hbm_mix_bench 610000000 > gpurun_out/hbm_mix_big.json 2>&1; cat gpurun_out/hbm_mix_big.json
