#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for cfg in "2 1" "4 1"; do set -- $cfg
BGR_TUNE_VEC=$1 BGR_TUNE_MINB=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_particles_program -s 14 -c 1 -f -o gpurun_out/prof2_v$1_m$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_v$1_m$2.log 2>&1
done
ls -la gpurun_out | tail -5
