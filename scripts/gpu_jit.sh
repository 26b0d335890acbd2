#!/bin/bash
# NVRTC-specialised generic program: parity tests on both kernels, then tick time vs the interpreter
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_component_presence.py tests/test_gpu_fuzz_requests.py tests/test_gpu_host_components.py tests/test_gpu_hierarchy.py tests/test_gpu_box_game.py tests/test_gpu_parity_particles.py -m gpu -x -q 2>&1 | tail -12
echo "pytest rc=$?"
for n in 100000 300000 1000000; do for v in "JIT=0" "JIT=2 BGR_TUNE_JIT_ITEM=512" "JIT=2 BGR_TUNE_JIT_ITEM=256" "JIT=2 BGR_TUNE_JIT_ITEM=128" "JIT=2 BGR_TUNE_JIT_ITEM=128 BGR_TUNE_JIT_ROWS=2" "JIT=2"; do
  echo "n=$n $v $(env BGR_JIT_VERBOSE=1 BGR_TUNE_$v timeout 300 python scripts/generic_world_bench.py $n 200 2>gpurun_out/jit_err.log | python -c "import sys,json; d=json.load(sys.stdin); print(d['presence_world_generic_program'], 'bundle', d['particles_bundle']['kernel_us_median'])") $(grep -c 'not specialised' gpurun_out/jit_err.log)"
done; done
