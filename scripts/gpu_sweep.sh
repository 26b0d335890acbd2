#!/bin/bash
# parity + tuning sweep of the fused kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for v in 1 2 4; do for m in 1 8; do for t in 1 0; do
  BGR_TUNE_VEC=$v BGR_TUNE_MINB=$m BGR_TUNE_PASSIVE_TMA=$t timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/sw_v${v}_m${m}_t${t}.log 2>&1
done; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); ok=True
            print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f ok=%s snap=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent'], {k:(round(v['save_us'],1),round(v['load_us'],1),round(v['frac_of_measured_hbm'],2)) for k,v in d.get('snapshot_save_restore',{}).items()}))
    if not ok: print(f, open(f).read()[-400:])
PY
