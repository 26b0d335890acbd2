#!/bin/bash
# first GPU session of round 2: parity suite, default bench with the timeline, synchronous-path sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 300 --warmup 5 --trace-out gpurun_out/r02a_timeline.csv > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.err
timeout 900 python scripts/sync_sweep.py > gpurun_out/r02a_sweep.jsonl 2> gpurun_out/r02a_sweep.err; echo "sweep rc=$?"
tail -3 gpurun_out/r02a_sweep.err
