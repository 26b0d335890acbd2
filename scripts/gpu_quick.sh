#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pf_*.json
BGR_TUNE_TILEDEP=9 timeout 400 python -m pytest tests/test_gpu_parity_particles.py -m gpu -x -q -k "pipelined_submits or catch_up" 2>&1 | tail -3
run() { name=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 60 python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/pf_$name.json 2>gpurun_out/pf_$name.err; echo "$name rc=$?"; }
for td in 0 5 9 11 13; do run 1m_td$td stress_1m_d8 1000 BGR_TUNE_TILEDEP=$td; done
for td in 5 9; do run 100k_td$td stress_100k_d8 2000 BGR_TUNE_TILEDEP=$td; run p2p_td$td p2p_1m_maxpred8 500 BGR_TUNE_TILEDEP=$td; run d16_td$td stress_1m_d16 500 BGR_TUNE_TILEDEP=$td; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/pf_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']), (d.get('catch_up_batch') or {}).get('ms_per_tick'))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-300:])
PY
