#!/bin/bash
# block size of the generic one-launch program (256 threads x 2 rows / 512 x 1) at three world sizes
for n in 100000 300000 1000000; do for b in 64 128; do
  echo "n=$n block=$b $(BGR_TUNE_GENERIC_BLOCK=$b timeout 300 python scripts/generic_world_bench.py $n 200 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print(d['presence_world_generic_program'])")"
done; done
BGR_TUNE_GENERIC_BLOCK=64 timeout 600 python -m pytest tests/test_gpu_component_presence.py tests/test_gpu_box_game.py tests/test_gpu_fuzz_requests.py -m gpu -x -q 2>&1 | tail -2
