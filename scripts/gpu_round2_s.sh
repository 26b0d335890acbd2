#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/fuzz_diag.py > gpurun_out/r02s_fuzz_diag.log 2>&1; echo "diag rc=$?"
cat gpurun_out/r02s_fuzz_diag.log | cut -c1-900
