#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sw2_*.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for lib in t512 t256; do for v in 2 4; do for m in 1 8; do for dyn in 0 1; do
  if [ $lib = t256 ]; then export BGR_LIBRARY=$PWD/bevy_ggrs_b200/libbevy_ggrs_b200_t256.so; else unset BGR_LIBRARY; fi
  BGR_TUNE_VEC=$v BGR_TUNE_MINB=$m BGR_TUNE_DYNAMIC=$dyn timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/sw2_${lib}_v${v}_m${m}_d${dyn}.log 2>&1
done; done; done; done
unset BGR_LIBRARY
BGR_TUNE_POLL=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/sw2_nopoll.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw2_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); ok=True
            print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f ok=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent']))
    if not ok: print(f, open(f).read()[-300:])
PY
