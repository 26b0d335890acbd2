#!/bin/bash
# the particles stress world itself on the NVRTC-specialised generic kernel (bundle off) next to the hand-written bundle
mkdir -p gpurun_out
SWEEP_EXTRA='[{"BGR_TUNE_BUNDLE":0,"BGR_TUNE_JIT":2,"BGR_TUNE_JIT_ITEM":512},{"BGR_TUNE_BUNDLE":0,"BGR_TUNE_JIT":2,"BGR_TUNE_JIT_ITEM":256},{"BGR_TUNE_BUNDLE":0,"BGR_TUNE_JIT":2,"BGR_TUNE_JIT_ITEM":128},{"BGR_TUNE_BUNDLE":0,"BGR_TUNE_JIT":2,"BGR_TUNE_JIT_ITEM":128,"BGR_TUNE_JIT_ROWS":2}]' \
  BGR_JIT_VERBOSE=1 timeout 900 python scripts/sync_sweep.py stress_100k_d8 stress_1m_d8 > gpurun_out/jit_particles_sweep.jsonl 2> gpurun_out/jit_particles_sweep.err
echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/jit_particles_sweep.jsonl'):
    d=json.loads(l)
    if 'error' in d: print(d); continue
    print(d['workload'], d['env'], 'sync %.1f us' % d['sync_us_per_tick'], 'kernel %.1f us' % d['kernel']['kernel_us_median'], 'pipelined %.1f us' % d['pipelined_us_per_tick'], 'fused', d.get('fused'))
PY
grep -c "not specialised" gpurun_out/jit_particles_sweep.err
