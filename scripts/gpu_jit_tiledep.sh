#!/bin/bash
# overlap of consecutive launches of the generated kernel (BGR_TUNE_JIT_TILEDEP): parity with 4 vectors in flight, then pipelined tick time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_component_presence.py -m gpu -x -q -k "pipelined_submits" 2>&1 | tail -6
for n in 100000 1000000; do for td in 0 1; do
  echo "n=$n tiledep=$td $(GENERIC_BENCH_PIPELINED=1 BGR_TUNE_JIT_TILEDEP=$td timeout 300 python scripts/generic_world_bench.py $n 200 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print(d['presence_world_generic_program'])")"
done; done
