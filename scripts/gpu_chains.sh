#!/bin/bash
# A/B of BGR_TUNE_CHAINS (intra-GPU multi-stream tile-range split)
mkdir -p gpurun_out; rm -f gpurun_out/ch_*.json gpurun_out/ch_*.err
run() { name=$1; wl=$2; steps=$3; shift 3; env "$@" timeout 120 python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ch_$name.json 2>gpurun_out/ch_$name.err; }
for c in 3 4 5; do run 1m_minb8_c$c stress_1m_d8 1000 BGR_TUNE_CHAINS=$c BGR_TUNE_MINB=8; done
for c in 5; do run 1m_c$c stress_1m_d8 1000 BGR_TUNE_CHAINS=$c; done
for c in 1 4 8; do run 10m_c$c stress_10m_d32 40 BGR_TUNE_CHAINS=$c; done
for c in 1 4; do run 1md16_c$c stress_1m_d16 500 BGR_TUNE_CHAINS=$c; done
for c in 1 4; do run p2p_c$c p2p_1m_maxpred8 500 BGR_TUNE_CHAINS=$c; done
run2() { name=$1; n=$2; shift 2; env "$@" timeout 120 python bench.py --entities $n --steps 1000 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/ch_$name.json 2>gpurun_out/ch_$name.err; }
for c in 1 2; do run2 500k_c$c 500000 BGR_TUNE_CHAINS=$c; done
for c in 1 4 8; do run2 2m_c$c 2000000 BGR_TUNE_CHAINS=$c; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ch_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    except Exception as e:
        print(f, "FAILED", open(f.replace('.json','.err')).read()[-300:])
PY
