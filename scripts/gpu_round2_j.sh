#!/bin/bash
mkdir -p gpurun_out
SWEEP_EXTRA='[{"BGR_TUNE_PASSIVE_EARLY": 1}, {"BGR_TUNE_PASSIVE_EARLY": 0}]' timeout 900 python scripts/sync_sweep.py > gpurun_out/r02j_sweep.jsonl 2> gpurun_out/r02j_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02j_sweep.jsonl | cut -c1-420
timeout 600 python scripts/generic_world_bench.py 100000 200 > gpurun_out/r02j_generic_world.json 2> gpurun_out/r02j_generic_world.err; echo "generic rc=$?"; cat gpurun_out/r02j_generic_world.json
timeout 600 python -m pytest tests/test_third_party_kats.py tests/test_gpu_component_presence.py tests/test_gpu_box_game.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench20.json 2> gpurun_out/r02j_bench20.err; echo "bench20 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench20.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['config']['timed_regions'], d['roofline']['sync'])
PY
