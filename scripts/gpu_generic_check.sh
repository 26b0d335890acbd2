#!/bin/bash
# generic one-launch program: parity tests that exercise it, then its tick time next to the bundle (scripts/generic_world_bench.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_box_game.py tests/test_gpu_component_presence.py tests/test_gpu_hierarchy.py tests/test_gpu_parity_particles.py tests/test_gpu_fuzz_requests.py tests/test_gpu_engine_edges.py -m gpu -x -q 2>&1 | tail -5
echo "pytest rc=$?"
for n in 100000 1000000; do
  timeout 300 python scripts/generic_world_bench.py $n 300 2>gpurun_out/generic_world_$n.err | tee gpurun_out/generic_world_$n.json
done
