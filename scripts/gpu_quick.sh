#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pf_*.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { name=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 120 python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/pf_$name.json 2>gpurun_out/pf_$name.err; }
for rep in 1 2; do
run 1m_p1_$rep stress_1m_d8 1000 BGR_TUNE_PREFETCH=1
run 1m_p0_$rep stress_1m_d8 1000 BGR_TUNE_PREFETCH=0
done
run 1m_p1_c4 stress_1m_d8 1000 BGR_TUNE_PREFETCH=1 BGR_TUNE_CHAINS=4
run 1m_p1_m8 stress_1m_d8 1000 BGR_TUNE_PREFETCH=1 BGR_TUNE_MINB=8
run d16_p1 stress_1m_d16 500 BGR_TUNE_PREFETCH=1
run d16_p0 stress_1m_d16 500 BGR_TUNE_PREFETCH=0
run p2p_p1 p2p_1m_maxpred8 500 BGR_TUNE_PREFETCH=1
run p2p_p0 p2p_1m_maxpred8 500 BGR_TUNE_PREFETCH=0
run 10m_p1 stress_10m_d32 40 BGR_TUNE_PREFETCH=1
run 10m_p0 stress_10m_d32 40 BGR_TUNE_PREFETCH=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/pf_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-300:])
PY
