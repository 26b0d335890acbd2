#!/bin/bash
# ncu full captures of the generic one-launch program on the presence world (scripts/generic_world_bench.py):
# the registration's own kernel (k_generic_jit, NVRTC) at 100k and 1M, and the interpreter kernel (BGR_TUNE_JIT=0) at 100k and 1M
mkdir -p gpurun_out
for n in 100000 1000000; do
  s=$([ $n = 100000 ] && echo "" || echo "_1m")
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_jit -s 20 -c 1 -f -o gpurun_out/r02_prof_generic_jit$s python scripts/generic_world_bench.py $n 24 > gpurun_out/r02_ncu_generic_jit$s.log 2>&1; echo "jit $n ncu rc=$?"
  BGR_TUNE_JIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_program -s 20 -c 1 -f -o gpurun_out/r02_prof_generic$s python scripts/generic_world_bench.py $n 24 > gpurun_out/r02_ncu_generic$s.log 2>&1; echo "interpreter $n ncu rc=$?"
  timeout 600 python scripts/generic_world_bench.py $n 200 > gpurun_out/r02_generic_world$s.json 2> gpurun_out/r02_generic_world$s.err; echo "bench $n rc=$?"
done
