#!/bin/bash
# ncu full capture of the generic one-launch program (presence world, 100k and 1M) + the bench JSON next to it
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_program -s 20 -c 2 -f -o gpurun_out/r02_prof_generic python scripts/generic_world_bench.py 100000 24 > gpurun_out/r02_ncu_generic.log 2>&1; echo "generic ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_program -s 20 -c 1 -f -o gpurun_out/r02_prof_generic_1m python scripts/generic_world_bench.py 1000000 24 > gpurun_out/r02_ncu_generic_1m.log 2>&1; echo "generic 1m ncu rc=$?"
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02_generic_world.json 2> gpurun_out/r02_generic_world.err; echo "generic bench rc=$?"
timeout 600 python scripts/generic_world_bench.py 1000000 200 > gpurun_out/r02_generic_world_1m.json 2> gpurun_out/r02_generic_world_1m.err; echo "generic bench 1m rc=$?"
