#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_generic_program -s 4 -c 2 -f -o gpurun_out/r02m_prof_generic python scripts/generic_world_bench.py 100000 12 > gpurun_out/r02m_ncu_generic.log 2>&1; echo "generic ncu rc=$?"
tail -3 gpurun_out/r02m_ncu_generic.log
ls -la gpurun_out/r02m_prof_generic.ncu-rep
