#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sw3_*.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for lib in t512 t256 t128; do for v in 2 4; do for bps in 0 3 2; do
  if [ $lib != t512 ]; then export BGR_LIBRARY=$PWD/bevy_ggrs_b200/libbevy_ggrs_b200_$lib.so; else unset BGR_LIBRARY; fi
  BGR_TUNE_VEC=$v BGR_TUNE_BPS=$bps timeout 300 python bench.py --steps 500 --warmup 5 --no-cpu-baseline --no-snapshot-bench > gpurun_out/sw3_${lib}_v${v}_bps${bps}.log 2>&1
done; done; done
unset BGR_LIBRARY
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw3_*.log')):
    ok=False
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); ok=True
            print(f, "value=%.0f ms=%.4f frac=%.3f e2e=%.0f ok=%s clocks=%s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['synctest_consistent'],d['clocks']))
    if not ok: print(f, open(f).read()[-300:])
PY
