#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 256" "2 256"; do set -- $cfg
BGR_TUNE_VEC=$1 BGR_TUNE_BLOCK=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_particles_program -s 14 -c 2 -f -o gpurun_out/prof_v$1_b$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v$1_b$2.log 2>&1
done
ls -la gpurun_out
